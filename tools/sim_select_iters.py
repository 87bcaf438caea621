"""Iteration counts of pivot strategies for the per-lane 'k-th largest within slack' select used by the
K1 epilogue compaction (warp-synchronous: a warp pays the MAX over its 32 lanes)."""
import numpy as np

def f2key(v):
    b = np.float32(v).view(np.uint32).astype(np.int64)
    return np.where(b & 0x80000000, (~b) & 0xFFFFFFFF, b | 0x80000000)

def key2f(k):
    k = np.int64(k)
    b = np.where(k & 0x80000000, k ^ 0x80000000, (~k) & 0xFFFFFFFF)
    return np.uint32(b).view(np.float32)

def iters(v, k, slack, strategy):
    v = np.asarray(v, np.float32)
    lo, hi = int(f2key(v.min())) - 1, int(f2key(v.max()))
    c_lo, c_hi = len(v), 0
    it = 0
    while hi - lo > 1:
        span = hi - lo
        if strategy == "current":
            step = max(1, span >> (2 if it < 2 else 1))
            mid = lo + step
        elif strategy == "interp":
            if it % 3 == 2:
                mid = lo + max(1, span >> 1)
            else:
                flo, fhi = float(key2f(max(lo, int(f2key(v.min()))))), float(key2f(hi))
                T = k + slack / 2
                frac = (c_lo - T) / max(1e-9, (c_lo - c_hi))
                fm = flo + (fhi - flo) * frac
                mid = int(f2key(np.float32(fm)))
                mid = min(max(mid, lo + 1), hi - 1)
        it += 1
        c = int((v > key2f(mid)).sum())
        if c >= k:
            lo, c_lo = mid, c
            if c <= k + slack:
                return it
        else:
            hi, c_hi = mid, c
    return it

rng = np.random.default_rng(1)
for dist in ("normal", "t2.5"):
    for strategy in ("current", "interp"):
        for slack in (8, 16):
            worst = []
            for trial in range(200):
                mx = 0
                for lane in range(32):
                    n = rng.integers(56, 81)
                    x = rng.standard_normal(3072) if dist == "normal" else rng.standard_t(2.5, 3072)
                    x = np.sort(x)[::-1][: n * 3]           # candidates: random subset of the upper tail
                    cand = rng.choice(x[: n * 2], n, replace=False)
                    mx = max(mx, iters(cand, 32, slack, strategy))
                worst.append(mx)
            print(f"{dist:7s} {strategy:8s} slack={slack}: warp iterations mean={np.mean(worst):.1f} max={np.max(worst)}")
