set -x
O=gpurun_out; T=r4a
timeout 600 python -W ignore::UserWarning -m pytest tests/test_gpu_encode_topk.py -m gpu -q -x > $O/${T}_pytest.log 2>&1; echo "rc=$?" >> $O/${T}_pytest.log
tail -n 25 $O/${T}_pytest.log
timeout 300 python tools/bench_k1_small.py 0 1 12 > $O/${T}_k1_small.txt 2>&1
timeout 300 python tools/profile_small_batch.py 128 > $O/${T}_small.txt 2>&1
cat $O/${T}_k1_small.txt $O/${T}_small.txt
