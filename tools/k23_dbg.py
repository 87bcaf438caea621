import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tools.bench_k23 as K
K.ALL_FIRED = True
for shape in ((75776, 384, 3072), (75776, 768, 6144), (37888, 1280, 40960)):
    for mode in (1, 2, 3, 0):
        t, m, _ = K.run(*shape, 32, mode, reps=8)
        print(f"all fired {shape}, mode={mode}: {t*1e3:.1f} us (min {m*1e3:.1f})", flush=True)
