#!/usr/bin/env python
"""K1 micro-benchmark: time wsae_encode_topk alone (CUDA events, L2-cold rotation of inputs).

    python tools/bench_k1.py [--d 384 --F 3072 --k 32 --batches 16384,65536] [--modes 0,1,2]

modes (experiments, wsae_debug_encode_mode): 0 = product kernel, 1 = GEMM pipeline only (epilogue
releases the accumulators unread), 2 = scan without compaction (results invalid; both need
--variant 1), 3 = scanner + selector kernel with the filter closed (pipeline + bare scan).
"""
import argparse
import ctypes
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
    ap.add_argument("--batches", default="16384,65536")
    ap.add_argument("--modes", default="0,1,2")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--counters", action="store_true")
    ap.add_argument("--variant", type=int, default=2, help="1 = one epilogue warp per quarter, 2 = scanner + selector")
    args = ap.parse_args()
    lib = _lib.load()
    lib.wsae_debug_encode_mode.argtypes = [ctypes.c_int]
    lib.wsae_debug_encode_variant.argtypes = [ctypes.c_int]
    lib.wsae_debug_encode_variant(args.variant)
    dev = "cuda"
    torch.manual_seed(0)
    w = torch.randn(args.F, args.d, device=dev) / args.d ** 0.5
    b = torch.zeros(args.F, device=dev)
    wp = ops.pack_encoder(w, b, 1)
    for B in [int(s) for s in args.batches.split(",")]:
        nrot = max(2, int(256e6 // (B * args.d * 2)) + 1)
        xs = [ops.pack_activations(torch.randn(B, args.d, device=dev), None, 1) for _ in range(nrot)]
        for mode in [int(s) for s in args.modes.split(",")]:
            lib.wsae_debug_encode_mode(mode)
            if mode >= 3:     # lives in the instrumented build of the scanner + selector kernel
                lib.wsae_debug_encode_counters.argtypes = [ctypes.c_void_p]
                dbuf = torch.zeros(148 * 8 * 8, dtype=torch.int64, device=dev)
                lib.wsae_debug_encode_counters(dbuf.data_ptr())
            for i in range(3):
                ops.encode_topk(xs[i % nrot], wp, B, args.F, args.d, 1, args.k)
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for i in range(args.iters):
                ops.encode_topk(xs[i % nrot], wp, B, args.F, args.d, 1, args.k)
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / args.iters
            if mode >= 3:
                torch.cuda.synchronize()
                dbuf.zero_()
                ops.encode_topk(xs[0], wp, B, args.F, args.d, 1, args.k)
                torch.cuda.synchronize()
                lib.wsae_debug_encode_counters(None)
                c3 = dbuf.view(148, 8, 8).double()[:, :4]
                it3 = B / 128 / 148
                print(f"  mode {mode} scanner: wait-tmem {c3[..., 2].mean() / it3:9.0f} cyc/item  total {c3[..., 3].mean() / it3:9.0f}")
            tf = 2.0 * B * args.d * args.F / ms / 1e9
            print(f"B={B} d={args.d} F={args.F} k={args.k} mode={mode}: {ms * 1e3:8.1f} us  "
                  f"{tf:7.1f} TFLOP/s (algorithmic)  {B / ms / 1e3:8.2f} Mrows/s", flush=True)
        lib.wsae_debug_encode_mode(0)
        if args.counters:     # scanner / selector wait breakdown of the v2 epilogue (one launch)
            lib.wsae_debug_encode_counters.argtypes = [ctypes.c_void_p]
            buf = torch.zeros(148 * 8 * 8, dtype=torch.int64, device=dev)
            lib.wsae_debug_encode_counters(buf.data_ptr())
            ops.encode_topk(xs[0], wp, B, args.F, args.d, 1, args.k)
            torch.cuda.synchronize()
            lib.wsae_debug_encode_counters(None)
            c = buf.view(148, 8, 8).double()
            items = B / 128 / 148
            sc, se = c[:, :4], c[:, 4:]
            print(f"  scanner : handoffs/item {sc[..., 0].mean() / items:6.1f}  wait-empty {sc[..., 1].mean() / items:9.0f} cyc/item"
                  f"  wait-tmem {sc[..., 2].mean() / items:9.0f}  total {sc[..., 3].mean() / items:9.0f}")
            print(f"  selector: selections/item {se[..., 0].mean() / items:6.1f}  wait-full {se[..., 1].mean() / items:9.0f} cyc/item"
                  f"  total {se[..., 3].mean() / items:9.0f}", flush=True)
            print(f"  selector sections (cyc/item): load {se[..., 2].mean() / items:8.0f}  select {se[..., 4].mean() / items:8.0f}"
                  f"  repack {se[..., 5].mean() / items:8.0f}  reload {se[..., 6].mean() / items:8.0f}  final {se[..., 7].mean() / items:8.0f}", flush=True)
        del xs


if __name__ == "__main__":
    main()
