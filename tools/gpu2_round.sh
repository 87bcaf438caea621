# bash tools/gpu2_round.sh TAG : 2-GPU checks (real NCCL): the DP parity test, 2-GPU bench
TAG=${1:-r2}
O=gpurun_out
set -x
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
timeout 900 python -W ignore::UserWarning -m pytest tests/test_parallel.py -m gpu -q -s > $O/${TAG}_pytest_dp.log 2>&1; echo "rc=$?" >> $O/${TAG}_pytest_dp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/${TAG}_bench_2gpu.json 2> $O/${TAG}_bench_2gpu.err; echo "rc=$?" >> $O/${TAG}_bench_2gpu.err
tail -5 $O/${TAG}_pytest_dp.log; tail -3 $O/${TAG}_bench_2gpu.err
