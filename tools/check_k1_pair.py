#!/usr/bin/env python
"""K1: the CTA-pair kernel (variant 3, tcgen05.mma.cta_group::2) against the single-CTA kernel
(variant 2) on the same packed operands: identical TopK sets / values, and both timed.

    python tools/check_k1_pair.py [--shapes 300x384x3072,75776x384x3072,...]
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whisper_sae_b200 import _lib, ops  # noqa: E402


def run(lib, variant, a, w, B, F, d, k, iters):
    lib.wsae_debug_encode_variant(variant)
    idx, val = ops.encode_topk(a[0], w, B, F, d, 1, k)
    torch.cuda.synchronize()
    for i in range(3):
        ops.encode_topk(a[i % len(a)], w, B, F, d, 1, k)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(iters):
        ops.encode_topk(a[i % len(a)], w, B, F, d, 1, k)
    t1.record()
    torch.cuda.synchronize()
    return idx, val, t0.elapsed_time(t1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="300x384x3072,4096x384x3072,75776x384x3072,128x384x3072,"
                                        "75776x768x6144,37888x1280x40960,1000x64x256")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    lib = _lib.load()
    ok_all = True
    for sh in args.shapes.split(","):
        B, d, F = (int(v) for v in sh.split("x"))
        k = min(32, F // 4)
        g = torch.Generator().manual_seed(B + d)
        wmat = (torch.randn(F, d, generator=g) / d ** 0.5).cuda()
        bias = (0.1 * torch.randn(F, generator=g)).cuda()
        w = ops.pack_encoder(wmat, bias, 1)
        nrot = max(2, int(256e6 // (B * d * 2)) + 1)
        nrot = min(nrot, 8)
        a = [ops.pack_activations(torch.randn(B, d, generator=g).cuda(), None, 1) for _ in range(nrot)]
        i2, v2, t2 = run(lib, 2, a, w, B, F, d, k, args.iters)
        i3, v3, t3 = run(lib, 3, a, w, B, F, d, k, args.iters)
        s2, o2 = torch.sort(i2, dim=1)
        s3, o3 = torch.sort(i3, dim=1)
        same_idx = bool(torch.equal(s2, s3))
        same_val = bool(torch.equal(v2.gather(1, o2), v3.gather(1, o3)))
        bad_rows = int((s2 != s3).any(1).sum())
        ok_all &= same_idx and same_val
        tf = 2.0 * B * d * F / 1e9
        print(f"B={B} d={d} F={F} k={k}: single-CTA {t2 * 1e3:8.1f} us ({tf / t2:7.1f} TF)  CTA-pair {t3 * 1e3:8.1f} us "
              f"({tf / t3:7.1f} TF)  idx equal {same_idx} (rows differing {bad_rows})  val equal {same_val}", flush=True)
    lib.wsae_debug_encode_variant(0)
    print("OK" if ok_all else "MISMATCH")
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
