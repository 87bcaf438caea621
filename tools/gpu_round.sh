# bash tools/gpu_round.sh TAG : full GPU test suite + smoke + bench (both arms) on one box
TAG=${1:-r2}
O=gpurun_out
set -x
timeout 1200 python -W ignore::UserWarning -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?" >> $O/${TAG}_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_ref.json 2> $O/${TAG}_ref.err
timeout 300 python tools/bench_k23.py > $O/${TAG}_k23.txt 2>&1
./tools/l2_gather_bench > $O/${TAG}_l2_gather.json 2>&1
./tools/l2_gather_bench 6144 768 75776 > $O/${TAG}_l2_gather_768.json 2>&1
tail -5 $O/${TAG}_pytest.log; cat $O/${TAG}_smoke.log | tail -3
