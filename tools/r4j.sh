set -x
O=gpurun_out; T=r4j
timeout 600 python -W ignore::UserWarning -m pytest tests/test_gpu_sparse.py -m gpu -q -x > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
tail -n 4 $O/${T}_pytest.log
for H in 0 1 3 2; do
  echo "== WSAE_K23_L2HINT=$H" >> $O/${T}_k23.txt
  WSAE_K23_L2HINT=$H timeout 300 python tools/bench_k23.py >> $O/${T}_k23.txt 2>&1
done
cat $O/${T}_k23.txt
