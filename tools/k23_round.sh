set -x
O=gpurun_out
T=${1:-r2u}
for PF in 0 1 2 3 5; do
  WSAE_K23_PF=$PF WSAE_K23_TC_CFG=32 timeout 300 python tools/bench_k23.py --all-fired --shapes 75776x384x3072,75776x768x6144,37888x1280x40960 > $O/${T}_k23_pf$PF.txt 2>&1
done
WSAE_K23_PF=2 timeout 600 python -m pytest tests/test_gpu_sparse.py -q -x -k fused > $O/${T}_pytest.log 2>&1; echo "rc=$?" >> $O/${T}_pytest.log
tail -n 4 $O/${T}_pytest.log; cut -c1-250 $O/${T}_k23_*.txt
