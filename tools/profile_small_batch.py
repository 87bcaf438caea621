#!/usr/bin/env python
"""Per-kernel CUDA-event spans of the eagerly launched train step at the YAML batch size (128 rows,
configs/tiny_default.yaml) next to the graphed step time.   python tools/profile_small_batch.py [rows]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from whisper_sae_b200 import ops  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    wl = bench.WORKLOADS["tiny"]
    dev = "cuda:0"
    batches = [bench.synth(rows, wl["d"], seed=s).to(dev) for s in range(4)]
    tr, _ = bench.make_trainer(wl, rows, dev, 0, cuda_graph="eager")
    for i in range(10):
        tr.train_step(batches[i % 4])
    prof = bench.kernel_profile(tr, batches, 40)
    step = prof.pop("_step_ms")
    print(f"rows={rows}: eager spans sum {step * 1e3:.1f} us per step")
    for n, e in sorted(prof.items(), key=lambda kv: -kv[1]["avg_ms"] * kv[1]["per_step"]):
        print(f"  {n:28s} {e['avg_ms'] * 1e3:7.1f} us x {e['per_step']:.0f}")
    tr2, _ = bench.make_trainer(wl, rows, dev, 0)
    for i in range(20):
        tr2.train_step(batches[i % 4])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 400
    for i in range(n):
        tr2.train_step(batches[i % 4])
    b.record()
    torch.cuda.synchronize()
    print(f"graphed step: {a.elapsed_time(b) / n * 1e3:.1f} us")


if __name__ == "__main__":
    main()
