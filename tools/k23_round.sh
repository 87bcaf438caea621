set -x
O=gpurun_out
timeout 300 python tools/bench_k23.py > $O/r2d_k23.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_module.py tests/test_gpu_fullsize.py -q -s > $O/r2d_pytest.log 2>&1; echo "rc=$?" >> $O/r2d_pytest.log
timeout 300 python bench.py --no-side-workloads --no-cpu-baseline > $O/r2d_bench.json 2> $O/r2d_bench.err
cat $O/r2d_k23.txt; tail -4 $O/r2d_pytest.log
