set -x
O=gpurun_out; T=r4c
timeout 1200 python -W ignore::UserWarning -m pytest tests/test_gpu_variants.py tests/test_gpu_fullsize.py -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
tail -n 30 $O/${T}_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
tail -n 5 $O/${T}_bench.err
