import os, sys, subprocess
from pathlib import Path
root = Path(__file__).resolve().parent.parent
for st in ("2", "3"):
    env = dict(os.environ, WSAE_K23_STAGES=st)
    out = subprocess.run([sys.executable, str(root / "tools" / "bench_k23.py"), "--shapes", "75776x384x3072,75776x768x6144,37888x1280x40960"],
                         env=env, capture_output=True, text=True)
    print(f"--- WSAE_K23_STAGES={st}\n{out.stdout}{out.stderr[-500:]}", flush=True)
