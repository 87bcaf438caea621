# bash tools/final_round.sh TAG : GPU suite + smoke + both bench arms + launch list at the YAML batch
TAG=${1:-r2d}
O=gpurun_out
set -x
timeout 1200 python -W ignore::UserWarning -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?" >> $O/${TAG}_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_ref.json 2> $O/${TAG}_ref.err
timeout 300 python bench.py --batch 128 --steps 50 --warmup 10 --value-only > $O/${TAG}_yaml_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_yaml_launches.csv \
    python bench.py --batch 128 --steps 50 --warmup 10 --value-only > $O/${TAG}_yaml_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'row_step|encode_dense' -s 20 -c 2 -f -o $O/${TAG}_yaml_top \
    python bench.py --batch 128 --steps 50 --warmup 10 --value-only > $O/${TAG}_yaml_ncu_full.log 2>&1
tail -4 $O/${TAG}_pytest.log; tail -2 $O/${TAG}_smoke.log; tail -2 $O/${TAG}_bench.err
