set -x
O=gpurun_out; T=r4t
timeout 1200 python -W ignore::UserWarning -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
tail -n 8 $O/${T}_pytest.log
for R in 128 256 512 1024; do
  WSAE_ROW_STEP_ROWS=1024 python tools/bench_replay_floor.py --batch $R --steps 400 >> $O/${T}_floor.txt 2>&1
  WSAE_ROW_STEP_ROWS=0 python tools/bench_replay_floor.py --batch $R --steps 400 >> $O/${T}_floor.txt 2>&1
done
WSAE_RAW_LAUNCH=0 python tools/bench_replay_floor.py --batch 128 --steps 400 >> $O/${T}_floor.txt 2>&1
cat $O/${T}_floor.txt
