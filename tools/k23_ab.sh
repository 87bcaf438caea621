set -x
O=gpurun_out
T=${1:-r3b}
for TC in 1 0 1 0; do
  for WL in tiny small-dp; do
    WSAE_K23_TC=$TC timeout 300 python bench.py --workload $WL --no-side-workloads --no-cpu-baseline --steps 20 --warmup 5 >> $O/${T}_ab_tc$TC.json 2>> $O/${T}_ab.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3b_ab_tc*.json')):
    for line in open(f):
        try: d=json.loads(line)
        except Exception: continue
        print(f, d['config']['workload'][:30], d['ms_per_step'], d.get('kernel_avg_ms',{}).get('wsae_decode_backward'))
PY
