#!/usr/bin/env python
"""K1 latency at the YAML batch sizes (64 / 128 rows) for several F-split counts, CUDA-graphed.

    python tools/bench_k1_small.py [nsplit ...]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whisper_sae_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
d, F, k = 384, 3072, 32
w = torch.randn(F, d, device=dev) / d ** 0.5
b = torch.zeros(F, device=dev)
wp = ops.pack_encoder(w, b, 1)
splits = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 4, 6, 12]   # 0 = default route (dense form up to 1024 rows)
for B in (64, 128, 1024):
    x = ops.pack_activations(torch.randn(B, d, device=dev), None, 1)
    for ns in splits:
        for _ in range(5):
            ops.encode_topk(x, wp, B, F, d, 1, k, nsplit=ns or None)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                ops.encode_topk(x, wp, B, F, d, 1, k, nsplit=ns or None)
        g.replay()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        g.replay()
        t1.record()
        torch.cuda.synchronize()
        print(f"B={B} nsplit={ns}: {t0.elapsed_time(t1) / 20 * 1e3:7.1f} us per call (graphed)")
