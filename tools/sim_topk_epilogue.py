"""Design-space simulation of the K1 epilogue's streaming TopK: how many warp-level
compactions does one work item trigger for a given list capacity / slack / column share?"""
import numpy as np, sys

def sim(F=3072, k=32, CAP=80, slack=8, chunk=16, share=1, rows=32, trials=20, seed=0):
    rng = np.random.default_rng(seed)
    comps, accepted = [], []
    for t in range(trials):
        x = rng.standard_normal((rows, F)).astype(np.float32)
        cols = F // share
        xs = x[:, :cols]
        tau = np.full(rows, -np.inf, np.float32)
        lists = [[] for _ in range(rows)]
        ncomp = 0; acc = 0
        for c0 in range(0, cols, chunk):
            blk = xs[:, c0:c0+chunk]
            for r in range(rows):
                sel = blk[r][blk[r] > tau[r]]
                acc += len(sel)
                lists[r].extend(sel.tolist())
            if max(len(l) for l in lists) > CAP - chunk:
                ncomp += 1
                for r in range(rows):
                    l = lists[r]
                    if len(l) > k + slack:
                        l.sort(reverse=True)
                        keep = k + slack // 2
                        tau[r] = l[keep]  # everything > l[keep] kept (keep elements)
                        lists[r] = l[:keep]
        comps.append(ncomp); accepted.append(acc / rows)
    return np.mean(comps), np.mean(accepted)

if __name__ == "__main__":
    for F in (3072, 40960):
        for CAP, slack, share in ((80, 8, 1), (64, 8, 2), (64, 4, 2), (96, 8, 2), (128, 8, 1), (160, 8, 1), (128, 16, 1), (96,8,1), (256, 16, 1)):
            c, a = sim(F=F, CAP=CAP, slack=slack, share=share, trials=6)
            print(f"F={F} CAP={CAP} slack={slack} share=1/{share}: compactions/item={c:.1f} accepted/row={a:.0f}")


def sim_hist(F=3072, k=32, CAP=72, bins=32, chunk=16, rows=32, trials=6, seed=0, dist="normal"):
    """Same stream, but compaction = one uniform-bin histogram pass (keep bins >= crossing bin)."""
    rng = np.random.default_rng(seed)
    comps, kept_after, noprog = [], [], 0
    for t in range(trials):
        if dist == "normal":
            x = rng.standard_normal((rows, F)).astype(np.float32)
        else:  # heavy tail
            x = rng.standard_t(2.5, (rows, F)).astype(np.float32)
        tau = np.full(rows, -np.inf, np.float32)
        lists = [[] for _ in range(rows)]
        ncomp = 0
        for c0 in range(0, F, chunk):
            blk = x[:, c0:c0+chunk]
            for r in range(rows):
                lists[r].extend(blk[r][blk[r] > tau[r]].tolist())
            if max(len(l) for l in lists) > CAP - chunk:
                ncomp += 1
                for r in range(rows):
                    l = np.array(lists[r], np.float32)
                    if len(l) <= k + 8:
                        continue
                    lo = tau[r] if np.isfinite(tau[r]) else l.min()
                    hi = l.max()
                    if hi <= lo:
                        noprog += 1; continue
                    b = np.minimum(bins - 1, ((l - lo) * (bins / (hi - lo))).astype(np.int64))
                    cnts = np.bincount(b, minlength=bins)
                    cum = np.cumsum(cnts[::-1])[::-1]          # cum[b] = count in bins >= b
                    bstar = np.max(np.nonzero(cum >= k)[0])
                    keep = b >= bstar
                    if keep.all():
                        noprog += 1; continue
                    tau[r] = l[~keep].max()
                    lists[r] = l[keep].tolist()
                    kept_after.append(keep.sum())
        comps.append(ncomp)
    return np.mean(comps), np.mean(kept_after), np.max(kept_after), noprog


if __name__ == "__main__":
    print("--- histogram compaction ---")
    for F in (3072, 40960):
        for dist in ("normal", "t2.5"):
            for CAP, bins in ((72, 32), (72, 16), (64, 32), (80, 32), (96, 32)):
                c, ka, km, nop = sim_hist(F=F, CAP=CAP, bins=bins, dist=dist, trials=4)
                print(f"F={F} {dist} CAP={CAP} bins={bins}: compactions/item={c:.1f} kept avg={ka:.1f} max={km} no-progress={nop}")
