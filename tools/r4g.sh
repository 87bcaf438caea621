set -x
O=gpurun_out; T=r4g
timeout 600 python -W ignore::UserWarning -m pytest tests/test_gpu_module.py tests/test_gpu_fullsize.py -m gpu -q -x > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
tail -n 5 $O/${T}_pytest.log
for R in 128 1024 4096; do
  WSAE_FORK_K4_ROWS=0 python tools/profile_small_batch.py $R 2>&1 | grep "graphed step" | sed "s/^/rows=$R k4-serial /" >> $O/${T}_small.txt
  python tools/profile_small_batch.py $R 2>&1 | grep "graphed step" | sed "s/^/rows=$R k4-forked /" >> $O/${T}_small.txt
done
cat $O/${T}_small.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
tail -n 5 $O/${T}_bench.err
