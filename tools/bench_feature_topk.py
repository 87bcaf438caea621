#!/usr/bin/env python
"""Micro-benchmark of the per-feature top-k tracker (wsae_feature_topk_update) at the bench shape,
next to the reference algorithm's CPU rate (oracle port of TopKTracker.update, bounded sample).

    python tools/bench_feature_topk.py [--rows 75776 --F 3072 --k 32 --K 20 --batches 6]

Algorithmic bytes per update: two passes over the (idx, val) code = 2 * 8 bytes per entry, plus
8 bytes written and read per surviving candidate (reported), against the measured HBM peak.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from whisper_sae_b200.analysis import TopKTracker  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=75776)
    ap.add_argument("--F", type=int, default=3072)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--K", type=int, default=20)
    ap.add_argument("--batches", type=int, default=6)
    ap.add_argument("--cpu-rows", type=int, default=512)
    args = ap.parse_args()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    codes = []
    for _ in range(args.batches):
        idx = torch.rand(args.rows, args.F, device=dev, generator=g).topk(args.k, dim=1).indices.to(torch.int32)
        val = torch.randn(args.rows, args.k, device=dev, generator=g).abs_()      # TopK codes are mostly > 0
        codes.append((idx.contiguous(), val.contiguous()))
    peaks = {}
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        peaks = json.loads(p.read_text())
    t = TopKTracker(args.F, args.K, device=dev)
    times = []
    for i, (idx, val) in enumerate(codes):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        t.update_sparse(idx, val, i * args.rows)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    n = args.rows * args.k
    steady = sorted(times[1:])[len(times[1:]) // 2]
    line = {
        "what": "feature_topk_update", "rows": args.rows, "entries": n, "F": args.F, "K": args.K,
        "first_batch_ms": times[0], "steady_ms": steady,
        "steady_rows_per_s": args.rows / steady * 1e3,
        "algorithmic_GBps_steady": 2 * 8 * n / steady / 1e6,
        "hbm_peak": peaks,
    }
    # CPU: the reference algorithm (oracle port) on a bounded sample of the same data
    from oracle.feature_topk_oracle import TrackerOracle
    o = TrackerOracle(args.F, args.K)
    idx, val = codes[0][0][: args.cpu_rows].cpu().numpy(), codes[0][1][: args.cpu_rows].cpu().numpy()
    t0 = time.perf_counter()
    o.update_sparse(idx, val, list(range(args.cpu_rows)))
    dt = time.perf_counter() - t0
    line["cpu_port_rows_per_s"] = args.cpu_rows / dt
    line["cpu_sample"] = f"{args.cpu_rows} rows of the first batch, oracle/feature_topk_oracle.py, 1 thread"
    print(json.dumps(line))


if __name__ == "__main__":
    main()
