set -x
O=gpurun_out; T=r4s
timeout 1200 python -W ignore::UserWarning -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
tail -n 8 $O/${T}_pytest.log
for R in 128 256 1024; do
  python tools/bench_replay_floor.py --batch $R --steps 400 >> $O/${T}_floor.txt 2>&1
done
WSAE_ROW_STEP_ROWS=0 python tools/bench_replay_floor.py --batch 128 --steps 400 >> $O/${T}_floor.txt 2>&1
WSAE_RAW_LAUNCH=0 python tools/bench_replay_floor.py --batch 128 --steps 400 >> $O/${T}_floor.txt 2>&1
python tools/bench_replay_floor.py --batch 75776 --steps 100 >> $O/${T}_floor.txt 2>&1
cat $O/${T}_floor.txt
python tools/profile_host_path.py > $O/${T}_host.txt 2>&1; head -30 $O/${T}_host.txt | cut -c1-150
