set -x
O=gpurun_out; T=r4i
for MODE in 0 1; do
  WSAE_DP_GRAPH=$MODE timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$MODE bench.py --gpus 2 --workload small-dp --steps 20 --warmup 5 --no-side-workloads --no-cpu-baseline > $O/${T}_small_graph$MODE.json 2> $O/${T}_small_graph$MODE.err; echo "rc=$?" >> $O/${T}_small_graph$MODE.err
  tail -2 $O/${T}_small_graph$MODE.err
done
nvidia-smi --query-gpu=index,utilization.gpu,memory.used --format=csv
