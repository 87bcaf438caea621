#!/usr/bin/env python
"""K23 (wsae_decode_backward) alone: the staged kernel (cp.async ring, default for d % 128 == 0), the
mma.sync variant and the general kernel on the same inputs - outputs compared, all timed with CUDA events.

    python tools/bench_k23.py [--shapes 75776x384x3072,75776x768x6144,37888x1280x40960]
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whisper_sae_b200 import _lib, ops  # noqa: E402


ALL_FIRED = False


def run(B, d, F, k, mode, reps=20):
    """mode: 0 = the library's per-shape choice, 1 = general (register staged), 2 = mma.sync dots, 3 = staged (cp.async ring), 4 = tensor form (both contractions on mma.sync)"""
    lib = _lib.load()
    lib.wsae_debug_decode_backward_general(mode)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, d, generator=g).cuda()
    w = (torch.randn(F, d, generator=g) / d ** 0.5).cuda().to(torch.bfloat16)
    b_dec = (0.1 * torch.randn(d, generator=g)).cuda()
    b_pre = (0.1 * torch.randn(d, generator=g)).cuda()
    idx = torch.randint(0, F, (B, k), generator=g, dtype=torch.int32).cuda()
    val = torch.randn(B, k, generator=g).cuda()
    if ALL_FIRED:      # what K1 hands over in training: the top-k pre-activations are positive, every entry fires
        val = val.abs() + 0.01
    last = torch.zeros(F, dtype=torch.int64, device="cuda")
    step = torch.zeros((), dtype=torch.int64, device="cuda")
    outs = None
    ts = []
    for it in range(reps + 3):
        resid_bf = torch.empty(B, d, dtype=torch.bfloat16, device="cuda")
        dpre = torch.empty(B, k, device="cuda")
        stats = torch.zeros(3, dtype=torch.int64, device="cuda")
        db_enc = torch.zeros(F, device="cuda")
        db_dec = torch.zeros(d, device="cuda")
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(2_000_000)
        a.record()
        ops.decode_backward(x, w, b_dec, b_pre, idx, val, None, 2.0 / (B * d), resid=None, resid_bf16=resid_bf,
                            stats=stats, last_activated=last, step_count=step, d_b_enc=db_enc, d_b_dec=db_dec,
                            dpre_val=dpre)
        b.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(a.elapsed_time(b))
        outs = (resid_bf.float(), dpre, stats.clone(), db_enc, db_dec)
    lib.wsae_debug_decode_backward_general(0)
    ts.sort()
    return ts[len(ts) // 2], ts[0], outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="75776x384x3072,75776x768x6144,37888x1280x40960,75776x1536x3072")
    ap.add_argument("--all-fired", action="store_true", help="all selected values positive (as in training)")
    args = ap.parse_args()
    global ALL_FIRED
    ALL_FIRED = args.all_fired
    for sh in args.shapes.split(","):
        B, d, F = (int(v) for v in sh.split("x"))
        k = 32
        t_gen, m_gen, o_gen = run(B, d, F, k, 1)
        t_mma, m_mma, _ = run(B, d, F, k, 2)
        t_fast, m_fast, o_staged = run(B, d, F, k, 3)
        t_tc, m_tc, o_fast = run(B, d, F, k, 4)
        t_auto, _, _ = run(B, d, F, k, 0)
        gather = B * k * d * 2
        errs = []
        for name, a, b in zip(("resid_bf16", "dpre", "stats", "db_enc", "db_dec"), o_fast, o_gen):
            if name == "stats":
                sse_a, sse_b = a[:1].view(torch.float64).item(), b[:1].view(torch.float64).item()
                errs.append(f"sse rel {abs(sse_a - sse_b) / abs(sse_b):.1e} l0 {'==' if a[1] == b[1] else '!='}")
            else:
                errs.append(f"{name} rel-L2 {((a - b).norm() / b.norm().clamp_min(1e-30)).item():.1e}")
        print(f"B={B} d={d} F={F}: general {t_gen * 1e3:.1f} us (min {m_gen * 1e3:.1f}), mma {t_mma * 1e3:.1f} us, "
              f"staged {t_fast * 1e3:.1f} us (min {m_fast * 1e3:.1f}), tensor form {t_tc * 1e3:.1f} us (min {m_tc * 1e3:.1f}) = "
              f"{gather / t_tc / 1e6:.0f} GB/s gathered; library choice {t_auto * 1e3:.1f} us; tensor form vs general: " + "; ".join(errs))


if __name__ == "__main__":
    main()
