set -x
O=gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'decode_backward_staged' -s 4 -c 1 -f -o $O/r2j_k23_staged python tools/bench_k23.py --shapes 75776x384x3072 > $O/r2j_ncu.log 2>&1
tail -3 $O/r2j_ncu.log
