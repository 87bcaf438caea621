#!/usr/bin/env python
"""Kernel-time table (torch.profiler, CUDA activities) of one hand-stepped variant step at B = 16384.
    python tools/profile_variants.py [crosscoder|transcoder|skip]"""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from whisper_sae_b200.sae import SkipTranscoder, TopKCrossLayerCrosscoder, TopKTranscoder, make_optimizer  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "crosscoder"
rows, d, F, k, L = 16384, 384, 3072, 32, 4
dev = "cuda:0"
acts = {li: bench.synth(rows, d, seed=50 + li).to(dev) for li in range(L)}
if kind == "crosscoder":
    m = TopKCrossLayerCrosscoder(d, L, F, k=k).to(dev)
    call = lambda: m(acts)
else:
    m = (SkipTranscoder if kind == "skip" else TopKTranscoder)(d, d, F, k=k).to(dev)
    call = lambda: m(acts[0], acts[1])
opt = make_optimizer(m)


def step():
    with torch.amp.autocast("cuda", dtype=torch.bfloat16):
        o = call()
    opt.zero_grad(set_to_none=False)
    o.loss.backward()
    opt.step()
    m.normalize_decoder_weights()


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
