# Round-2 profile set: GPU tests, bench line, ncu launch lists + `--set full` captures of the three
# big kernels at the three widths (tiny / small / large-v3).  Usage: bash tools/profile_r2.sh TAG [quick]
TAG=${1:-r2}
QUICK=${2:-}
O=gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt
if [ -z "$QUICK" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
fi
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
for WL in tiny small-dp large-dp; do
  timeout 300 python bench.py --workload $WL --steps 5 --warmup 3 --value-only > $O/${TAG}_${WL}_plain.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${TAG}_${WL}_launches.csv \
      python bench.py --workload $WL --steps 5 --warmup 3 --value-only > $O/${TAG}_${WL}_ncu.log 2>&1
  timeout 900 ncu --set full --import-source on --clock-control none \
      -k regex:'encode_topk2|encode_topk3|decode_backward|wgrad_gemm_kernel' -s 12 -c 4 -f -o $O/${TAG}_${WL}_top \
      python bench.py --workload $WL --steps 5 --warmup 3 --value-only > $O/${TAG}_${WL}_ncu_full.log 2>&1
done
ls -la $O | tail -30
