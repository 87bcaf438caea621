#!/usr/bin/env python
"""GPU floor of the graphed train step: the captured graph replayed back to back with no host work
in between (no control-block refresh, no metrics poll, no LR scheduler), next to the full
`SAETrainer.train_step` loop on the same trainer.  The difference is what the host path still costs.

    python tools/bench_replay_floor.py --batch 128 --steps 300
"""
import argparse
import json
import sys
import tempfile
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=300)
    args = ap.parse_args()
    from whisper_sae_b200.config import TrainingConfig
    from whisper_sae_b200.sae import SAETrainer, TopKSAE

    d, F, k, B = 384, 3072, 32, args.batch
    torch.manual_seed(42)
    sae = TopKSAE(d, F, k=k)
    cfg = TrainingConfig(batch_size=B, use_amp=True, num_workers=0)
    tr = SAETrainer(sae, cfg, device="cuda:0", run_dir=Path(tempfile.mkdtemp()))
    tr.setup_scheduler(100_000)
    g = torch.Generator().manual_seed(1)
    xs = [torch.randn(B, d, generator=g).cuda() for _ in range(8)]
    for i in range(20):
        tr.train_step(xs[i % 8])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(args.steps):
        tr.train_step(xs[i % 8])
    b.record()
    host_s = time.perf_counter() - t0          # host time to ISSUE the loop (the last step still runs)
    torch.cuda.synchronize()
    full_ms = a.elapsed_time(b) / args.steps
    gs = tr._graphs[B]
    a.record()
    for i in range(args.steps):
        gs.graph.replay()
    b.record()
    torch.cuda.synchronize()
    floor_ms = a.elapsed_time(b) / args.steps
    print(json.dumps({"batch_rows": B, "steps": args.steps, "train_step_ms": full_ms,
                      "replay_only_ms": floor_ms, "host_issue_ms_per_step": 1e3 * host_s / args.steps}))


if __name__ == "__main__":
    main()
