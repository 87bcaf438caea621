set -x
O=gpurun_out; T=r4h
timeout 900 python -W ignore::UserWarning -m pytest tests/test_parallel.py -m gpu -q -s > $O/${T}_pytest_dp.log 2>&1; echo "rc=$?" >> $O/${T}_pytest_dp.log
tail -12 $O/${T}_pytest_dp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/${T}_bench_2gpu.json 2> $O/${T}_bench_2gpu.err; echo "rc=$?" >> $O/${T}_bench_2gpu.err
WSAE_DP_OPERANDS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/${T}_bench_2gpu_fp32gather.json 2> $O/${T}_bench_2gpu_fp32gather.err
tail -3 $O/${T}_bench_2gpu.err
