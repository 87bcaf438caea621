set -x
O=gpurun_out; T=r4e
timeout 1200 python -W ignore::UserWarning -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
tail -n 30 $O/${T}_pytest.log
python tools/profile_variants.py crosscoder > $O/${T}_cc.txt 2>&1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
tail -n 5 $O/${T}_bench.err
