set -x
O=gpurun_out; T=r4b
timeout 1200 python -W ignore::UserWarning -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
tail -n 5 $O/${T}_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
tail -n 5 $O/${T}_bench.err
