# bash tools/gpu8_round.sh TAG N : N-GPU bench (real NCCL), both arms' launch form
TAG=${1:-r2}
N=${2:-8}
O=gpurun_out
set -x
nvidia-smi topo -m > $O/${TAG}_topo_${N}.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err; echo "rc=$?" >> $O/${TAG}_bench_${N}gpu.err
tail -3 $O/${TAG}_bench_${N}gpu.err
