set -x
O=gpurun_out; T=r4n
timeout 600 python -W ignore::UserWarning -m pytest tests/test_gpu_sparse.py tests/test_gpu_encode_topk.py -m gpu -q -x -k "row_step or dense" > $O/${T}_pytest_a.log 2>&1; echo "rc=$?" >> $O/${T}_pytest_a.log
tail -n 15 $O/${T}_pytest_a.log
timeout 900 python -W ignore::UserWarning -m pytest tests/test_gpu_module.py -m gpu -q -x > $O/${T}_pytest_b.log 2>&1; echo "rc=$?" >> $O/${T}_pytest_b.log
tail -n 15 $O/${T}_pytest_b.log
python tools/profile_small_batch.py 128 > $O/${T}_small.txt 2>&1
WSAE_ROW_STEP_ROWS=0 python tools/profile_small_batch.py 128 >> $O/${T}_small.txt 2>&1
python tools/profile_small_batch.py 256 >> $O/${T}_small.txt 2>&1
WSAE_ROW_STEP_ROWS=0 python tools/profile_small_batch.py 256 >> $O/${T}_small.txt 2>&1
cat $O/${T}_small.txt
