set -x
python bench.py > gpurun_out/r1v12_bench.json 2> gpurun_out/r1v12_bench.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r1v12_ref.json 2> gpurun_out/r1v12_ref.err
python bench.py --steps 5 --warmup 3 --value-only > gpurun_out/r1v12_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1v12_launches.csv python bench.py --steps 5 --warmup 3 --value-only > gpurun_out/r1v12_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'encode_topk2|decode_backward_kernel|wgrad_gemm' -s 12 -c 4 -f -o gpurun_out/r1v12_top python bench.py --steps 5 --warmup 3 --value-only > gpurun_out/r1v12_ncu_full.log 2>&1
python tools/bench_feature_topk.py > gpurun_out/r1v12_tracker.json 2> gpurun_out/r1v12_tracker.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ftk_ -c 40 --csv --log-file gpurun_out/r1v12_tracker_launches.csv python tools/bench_feature_topk.py > /dev/null 2>&1
python tools/bench_k1.py --batches 75776 --modes 0,3,4 --counters > gpurun_out/r1v12_k1.log 2>&1
python tools/bench_k1.py --batches 75776 --modes 1 --variant 1 >> gpurun_out/r1v12_k1.log 2>&1
ls -la gpurun_out | tail -12
