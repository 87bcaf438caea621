#!/usr/bin/env python
"""K4 micro-benchmark: time wsae_wgrad_gemm alone (CUDA events, rotating inputs) for several cluster
sizes (TMA multicast of the R tiles), and check the result against a dense fp32 product.

    python tools/bench_k4.py [--d 384 --F 3072 --k 32 --rows 75776 --clusters 1,2,4]
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whisper_sae_b200 import _lib, ops  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, default=384)
    ap.add_argument("--F", type=int, default=3072)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--rows", type=int, default=75776)
    ap.add_argument("--clusters", default="1,2,4")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    lib = _lib.load()
    dev = "cuda"
    B, d, F, k = args.rows, args.d, args.F, args.k
    g = torch.Generator(device=dev).manual_seed(0)
    nrot = 4
    sets = []
    for _ in range(nrot):
        idx = torch.rand(B, F, device=dev, generator=g).topk(k, dim=1).indices.to(torch.int32).contiguous()
        val = torch.rand(B, k, device=dev, generator=g) + 0.1
        dpre = torch.randn(B, k, device=dev, generator=g)
        r = torch.randn(B, d, device=dev, generator=g).to(torch.bfloat16).contiguous()
        sets.append((ops.bucket_by_tile(idx, val, dpre, F), r, idx, val))
    out = torch.zeros(F, d, device=dev)
    ref = None
    for c in [int(s) for s in args.clusters.split(",")]:
        lib.wsae_debug_wgrad_cluster(c)
        out.zero_()
        bk, r, idx, val = sets[0]
        ops.wgrad_gemm_(out, r, B, d, bk, bk.act, None, 1.0)
        torch.cuda.synchronize()
        if ref is None:      # dense check on a slice of features (bf16-rounded activations, fp32 accumulate)
            nb = 8192
            dense = torch.zeros(nb, F, device=dev)
            dense.scatter_(1, idx[:nb].long(), val[:nb].to(torch.bfloat16).float())
            ref = (dense.t() @ r[:nb].float())
            out2 = torch.zeros(F, d, device=dev)
            bk2 = ops.bucket_by_tile(idx[:nb].contiguous(), val[:nb].contiguous(), val[:nb].contiguous(), F)
            ops.wgrad_gemm_(out2, r[:nb].contiguous(), nb, d, bk2, bk2.act, None, 1.0)
            err = (out2 - ref).abs().max().item() / ref.abs().max().item()
            print(f"cluster {c}: max rel err vs dense fp32 on {nb} rows: {err:.2e}")
            first = out.clone()
        else:
            print(f"cluster {c}: max |diff| vs first cluster size: {(out - first).abs().max().item():.3e} "
                  f"(scale {first.abs().max().item():.3e}; split-K order differs)")
        for i in range(3):
            bk, r, _, _ = sets[i % nrot]
            ops.wgrad_gemm_(out, r, B, d, bk, bk.act, None, 1.0)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(args.iters):
            bk, r, _, _ = sets[i % nrot]
            ops.wgrad_gemm_(out, r, B, d, bk, bk.act, None, 1.0)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / args.iters
        print(f"B={B} d={d} F={F} cluster<={c}: {ms * 1e3:8.1f} us  {2.0 * B * d * F / ms / 1e9:7.1f} TFLOP/s", flush=True)
    lib.wsae_debug_wgrad_cluster(2)


if __name__ == "__main__":
    main()
