# ncu --set full of the final K23 kernel (L2 eviction hints) at the three widths
O=gpurun_out
for WL in tiny small-dp large-dp; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:'decode_backward' -s 4 -c 1 -f -o $O/r2d_${WL}_k23 \
      python bench.py --workload $WL --steps 5 --warmup 3 --value-only > $O/r2d_${WL}_k23_ncu.log 2>&1
done
ls -la $O/r2d_*_k23.ncu-rep
