set -x
O=gpurun_out
T=${1:-r2v}
MODE=${2:-4}
cat > /tmp/k23_one.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
sys.argv = ['x']
import importlib.util
spec = importlib.util.spec_from_file_location('bk', 'tools/bench_k23.py'); bk = importlib.util.module_from_spec(spec); spec.loader.exec_module(bk)
bk.ALL_FIRED = True
import os
mode = int(os.environ.get('K23_MODE', '4'))
B, d, F = (int(v) for v in os.environ.get('K23_SHAPE', '75776x384x3072').split('x'))
print(bk.run(B, d, F, 32, mode, reps=3)[:2])
PY
K23_MODE=$MODE timeout 600 ncu --set full --import-source on --clock-control none -k regex:'decode_backward' -s 4 -c 1 -f -o $O/${T}_k23 python /tmp/k23_one.py > $O/${T}_ncu.log 2>&1
tail -3 $O/${T}_ncu.log
