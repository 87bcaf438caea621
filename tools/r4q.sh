O=gpurun_out; T=r4q
for R in 128 256; do
  python tools/bench_replay_floor.py --batch $R --steps 400 >> $O/${T}_floor.txt 2>&1
  WSAE_ROW_STEP_ROWS=0 python tools/bench_replay_floor.py --batch $R --steps 400 >> $O/${T}_floor.txt 2>&1
done
cat $O/${T}_floor.txt
