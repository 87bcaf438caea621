"""K2 / K3 / K5 parity through the C ABI against the CPU oracle."""

import math

import pytest
import torch

from oracle import topk_sae_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from whisper_sae_b200 import ops
    return ops


def _case(B, d, F, k, seed=0, b_pre_scale=0.05):
    torch.manual_seed(seed)
    state = O.init_state(d, F)
    state["b_pre"] = torch.randn(d) * b_pre_scale
    state["decoder.bias"] = torch.randn(d) * 0.02
    x = O.synthetic_activations(B, d, seed=seed + 100)
    fwd = O.forward(state, x, k, training=False)
    return state, x, fwd


def _dev(t):
    return t.cuda().contiguous()


SHAPES = [(1, 32, 128, 4), (33, 64, 256, 8), (64, 384, 3072, 32), (257, 768, 1024, 32),
          (40, 1280, 2048, 32), (16, 64, 128, 64)]


@pytest.mark.parametrize("B,d,F,k", SHAPES)
@pytest.mark.parametrize("wdtype", ["fp32", "bf16"])
def test_decode_mse(B, d, F, k, wdtype):
    ops = _ops()
    state, x, fwd = _case(B, d, F, k, seed=B + d)
    quant = "bf16" if wdtype == "bf16" else None
    # oracle decode with the same decoder precision (encoder side identical: same idx/val fed in)
    state_q = dict(state)
    ref = O.forward(state_q, x, k, training=False)
    w_decT = state["decoder.weight"].t().contiguous()
    h = torch.relu(ref.val)
    if quant:      # bf16 shadow => bf16 x bf16 products (activation rounded too), fp32 accumulation
        w_used = w_decT.to(torch.bfloat16)
        rows = w_used.float()[ref.idx]
        h = h.to(torch.bfloat16).float()
    else:
        w_used = w_decT
        rows = w_decT[ref.idx]
    recon = (h.unsqueeze(-1) * rows).sum(1) + state["decoder.bias"] + state["b_pre"]
    resid_ref = recon - x
    stats = torch.zeros(3, dtype=torch.int64, device="cuda")
    last = torch.zeros(F, dtype=torch.int64, device="cuda")
    step = torch.tensor(6, dtype=torch.int64, device="cuda")
    resid, rec = ops.decode_mse(_dev(x), _dev(w_used), _dev(state["decoder.bias"]), _dev(state["b_pre"]),
                                _dev(ref.idx.to(torch.int32)), _dev(ref.val), want_resid=True,
                                want_recon=True, stats=stats, last_activated=last, step_count=step)
    torch.cuda.synchronize()
    torch.testing.assert_close(resid.cpu(), resid_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rec.cpu(), recon, rtol=1e-5, atol=1e-6)
    raw = stats.cpu()
    sse = raw[:1].view(torch.float64).item()
    assert sse == pytest.approx((resid_ref.double() ** 2).sum().item(), rel=1e-6)
    assert raw[1].item() == int((ref.val > 0).sum())            # L0 numerator: exact integer
    want_last = torch.zeros(F, dtype=torch.int64)
    want_last[ref.idx[ref.val > 0]] = 7                          # stamp = step_count + 1
    assert torch.equal(last.cpu(), want_last)                    # bit-exact
    assert step.item() == 6                                      # K2 itself does not bump


@pytest.mark.parametrize("B,d,F,k", SHAPES)
def test_backward_sparse_fp32(B, d, F, k):
    ops = _ops()
    state, x, fwd = _case(B, d, F, k, seed=2 * B + d)
    grad_out = 3.0
    ref = O.backward(state, x, fwd, grad_out=grad_out)
    dev = "cuda"
    w_decT = _dev(state["decoder.weight"].t())
    dWe = torch.zeros(F, d, device=dev)
    dWd = torch.zeros(F, d, device=dev)
    dbe = torch.zeros(F, device=dev)
    dbd = torch.zeros(d, device=dev)
    dpre = torch.empty(B, k, device=dev)
    go = torch.tensor(grad_out, device=dev)
    coef = 2.0 / (B * d)
    idx = _dev(fwd.idx.to(torch.int32))
    ops.backward_sparse(_dev(fwd.resid), _dev(x), _dev(state["b_pre"]), w_decT, idx, _dev(fwd.val), go,
                        coef, d_w_enc=dWe, d_w_decT=dWd, d_b_enc=dbe, d_b_dec=dbd, dpre_val=dpre)
    dbp = ops.bpre_grad(dbd, dbe, _dev(state["encoder.weight"]))
    dx = ops.input_grad(_dev(fwd.resid), _dev(state["encoder.weight"]), idx, dpre, go, coef, True)
    torch.cuda.synchronize()

    def close(a, b, name):
        scale = b.abs().max().item() + 1e-30
        torch.testing.assert_close(a.cpu(), b, rtol=2e-5, atol=2e-6 * scale, msg=lambda m: f"{name}: {m}")

    close(dpre, ref["dpre_val"], "dpre")
    close(dWe, ref["encoder.weight"], "dW_enc")
    close(dWd.t(), ref["decoder.weight"], "dW_dec")
    close(dbe, ref["encoder.bias"], "db_enc")
    close(dbd, ref["decoder.bias"], "db_dec")
    close(dbp, ref["b_pre"], "db_pre")
    close(dx, ref["dx"], "dx")


FUSED_SHAPES = [(1, 32, 128, 4), (33, 64, 256, 8), (64, 384, 3072, 32), (257, 768, 1024, 32),
                (40, 1280, 2048, 32), (300, 392, 512, 17), (70, 128, 512, 17), (130, 256, 300, 5),
                (1000, 1536, 640, 32)]


def test_decode_backward_fused_rejects_fp32_decoder():
    ops = _ops()
    assert not ops.decode_backward_supported(384, 32, False)
    assert not ops.decode_backward_supported(384, 64, True)


@pytest.mark.parametrize("B,d,F,k", FUSED_SHAPES)
@pytest.mark.parametrize("wdtype", ["bf16"])
def test_decode_backward_fused(B, d, F, k, wdtype):
    """K23 (one pass over the gathered decoder rows) against the oracle forward + backward.
    The dot products use the bf16-rounded residual (the rounding K4 consumes), so dv / db_enc are
    compared at bf16 tolerance; everything on the forward side stays at fp32 tolerance."""
    ops = _ops()
    state, x, fwd = _case(B, d, F, k, seed=3 * B + d)
    if wdtype == "bf16":   # the decoder the kernel reads is the bf16 shadow: same rounding in the oracle
        state = dict(state)
        state["decoder.weight"] = state["decoder.weight"].to(torch.bfloat16).float()
        fwd = O.forward(state, x, k, training=False)
    # force a few selected values non-positive: relu-masked entries must contribute nothing
    val = fwd.val.clone()
    val[::3, 0] = -val[::3, 0].abs()
    h_bf = torch.relu(val).to(torch.bfloat16).float()          # bf16 x bf16 decode products
    recon = (h_bf.unsqueeze(-1) * state["decoder.weight"].t()[fwd.idx]).sum(1) \
        + state["decoder.bias"] + state["b_pre"]
    resid_ref = recon - x
    grad_out = 0.5
    coef = 2.0 / (B * d)
    g = resid_ref * (coef * grad_out)
    dv_ref = (g.unsqueeze(1) * state["decoder.weight"].t()[fwd.idx]).sum(-1) * (val > 0)
    dbe_ref = torch.zeros(F).index_add_(0, fwd.idx.reshape(-1), dv_ref.reshape(-1))
    dev = "cuda"
    w_decT = state["decoder.weight"].t().contiguous()
    w_used = _dev(w_decT.to(torch.bfloat16) if wdtype == "bf16" else w_decT)
    stats = torch.zeros(3, dtype=torch.int64, device=dev)
    last = torch.zeros(F, dtype=torch.int64, device=dev)
    step = torch.tensor(41, dtype=torch.int64, device=dev)
    resid = torch.empty(B, d, device=dev)
    resid_bf = torch.empty(B, d, dtype=torch.bfloat16, device=dev)
    dbe = torch.zeros(F, device=dev)
    dbd = torch.zeros(d, device=dev)
    dpre = torch.empty(B, k, device=dev)
    assert ops.decode_backward_supported(d, k, wdtype == "bf16")
    ops.decode_backward(_dev(x), w_used, _dev(state["decoder.bias"]), _dev(state["b_pre"]),
                        _dev(fwd.idx.to(torch.int32)), _dev(val), torch.tensor(grad_out, device=dev), coef,
                        resid=resid, resid_bf16=resid_bf, stats=stats, last_activated=last,
                        step_count=step, d_b_enc=dbe, d_b_dec=dbd, dpre_val=dpre)
    torch.cuda.synchronize()
    torch.testing.assert_close(resid.cpu(), resid_ref, rtol=1e-5, atol=2e-6)
    assert torch.equal(resid_bf.cpu(), resid.cpu().to(torch.bfloat16))
    raw = stats.cpu()
    assert raw[:1].view(torch.float64).item() == pytest.approx((resid_ref.double() ** 2).sum().item(), rel=1e-6)
    assert raw[1].item() == int((val > 0).sum())
    want_last = torch.zeros(F, dtype=torch.int64)
    want_last[fwd.idx[val > 0]] = 42
    assert torch.equal(last.cpu(), want_last)

    def close(a, b, name, rtol=2e-5, arel=2e-6):
        scale = b.abs().max().item() + 1e-30
        torch.testing.assert_close(a.cpu(), b, rtol=rtol, atol=arel * scale, msg=lambda m: f"{name}: {m}")

    close(dpre, dv_ref, "dpre", rtol=2e-2, arel=4e-3)
    close(dbe, dbe_ref, "db_enc", rtol=2e-2, arel=4e-3)
    close(dbd, g.sum(0), "db_dec")
    # and exactly the value the bf16-rounded residual implies (fp32 accumulation of bf16 products)
    g_bf = resid_bf.float().cpu() * (coef * grad_out)
    dv_bf = (g_bf.unsqueeze(1) * state["decoder.weight"].t()[fwd.idx]).sum(-1) * (val > 0)
    close(dpre, dv_bf, "dpre vs bf16-residual reference", rtol=1e-4, arel=1e-5)


def test_renorm_and_shadow():
    ops = _ops()
    torch.manual_seed(0)
    F, d = 1000, 384
    w = torch.randn(F, d) * torch.rand(F, 1) * 3
    w[5] = 0.0                      # zero row hits the 1e-12 clamp (SkipTranscoder case, §8 a15)
    ref = torch.nn.functional.normalize(w.t(), dim=0).t()
    wg = w.cuda()
    shadow = torch.empty(F, d, dtype=torch.bfloat16, device="cuda")
    ops.renorm_decoder_(wg, 1e-12, shadow)
    torch.cuda.synchronize()
    torch.testing.assert_close(wg.cpu(), ref, rtol=1e-6, atol=1e-7)
    assert (wg[5] == 0).all()
    assert torch.equal(shadow.cpu(), wg.cpu().to(torch.bfloat16))
    torch.testing.assert_close(wg.norm(dim=1).cpu()[torch.arange(F) != 5], torch.ones(F - 1), atol=1e-5, rtol=0)


def test_counters_bit_exact():
    ops = _ops()
    F = 40960
    g = torch.Generator().manual_seed(1)
    last = torch.randint(0, 5000, (F,), generator=g, dtype=torch.int64)
    step = torch.tensor(5200, dtype=torch.int64)
    for thr in (0, 1000, 10_000):
        lg, sg = last.cuda(), step.clone().cuda()
        dead = torch.zeros(1, dtype=torch.int64, device="cuda")
        ops.counters_update(lg, sg, thr, True, dead)
        torch.cuda.synchronize()
        assert sg.item() == 5201
        assert dead.item() == int(((5201 - last) > thr).sum())
        ops.counters_update(lg, sg, thr, False, dead)
        assert sg.item() == 5201 and dead.item() == int(((5201 - last) > thr).sum())


def test_densify_cast_sumsq():
    ops = _ops()
    state, x, fwd = _case(50, 64, 256, 8, seed=4)
    h = ops.densify_hidden(fwd.idx.to(torch.int32).cuda(), fwd.val.cuda(), 256)
    assert torch.equal(h.cpu(), O.dense_hidden(fwd, 256))
    t = torch.randn(1001, 37).cuda()
    assert torch.equal(ops.cast_bf16(t).cpu(), t.cpu().to(torch.bfloat16))
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    ops.sumsq_(t, acc)
    assert acc.item() == pytest.approx((t.double() ** 2).sum().item(), rel=1e-6)


def test_fused_adamw_matches_torch():
    ops = _ops()
    torch.manual_seed(0)
    shapes = [(384,), (3072, 384), (3072,), (3072, 384), (384,)]
    ps = [torch.randn(s) for s in shapes]
    ref_p = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.AdamW(ref_p, lr=1e-3, weight_decay=0.01)
    gp = [p.clone().cuda() for p in ps]
    m = [torch.zeros_like(p) for p in gp]
    v = [torch.zeros_like(p) for p in gp]
    hyper = torch.empty(8, device="cuda")
    for step in range(1, 4):
        grads = [torch.randn(s) * (10.0 if step == 2 else 0.01) for s in shapes]
        for p, g in zip(ref_p, grads):
            p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_(ref_p, 1.0)
        opt.step()
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        gg = [g.cuda() for g in grads]
        for g in gg:
            ops.sumsq_(g, ss)
        hyper.copy_(torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.01, 1 - 0.9 ** step,
                                  math.sqrt(1 - 0.999 ** step), 1.0]))
        for p, g, mm, vv in zip(gp, gg, m, v):
            ops.fused_adamw_(p, g, mm, vv, hyper, ss)
        torch.cuda.synchronize()
        for p, r in zip(gp, ref_p):
            torch.testing.assert_close(p.cpu(), r.detach(), rtol=2e-6, atol=1e-7)


def test_adamw_multi_matches_torch_with_renorm():
    """One launch: clip + AdamW over 5 tensors + unit-norm rows of the (feature-major) decoder."""
    ops = _ops()
    torch.manual_seed(1)
    F, d = 1000, 392
    shapes = [(d,), (F, d), (F,), (F, d), (d,)]
    ps = [torch.randn(s) for s in shapes]
    ref_p = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.AdamW(ref_p, lr=1e-3, weight_decay=0.01)
    gp = [p.clone().cuda() for p in ps]
    m = [torch.zeros_like(p) for p in gp]
    v = [torch.zeros_like(p) for p in gp]
    hyper = torch.empty(8, device="cuda")
    for step in range(1, 4):
        grads = [torch.randn(s) * (10.0 if step == 2 else 0.01) for s in shapes]
        for p, g in zip(ref_p, grads):
            p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_(ref_p, 1.0)
        opt.step()
        with torch.no_grad():    # decoder (index 3) stored feature-major: unit rows == unit columns of W_dec
            ref_p[3].copy_(torch.nn.functional.normalize(ref_p[3], dim=1))
        gg = [g.cuda() for g in grads]
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        for g in gg:
            ops.sumsq_(g, ss)
        hyper.copy_(torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.01, 1 - 0.9 ** step,
                                  math.sqrt(1 - 0.999 ** step), 1.0]))
        entries = [(p, g, mm, vv, d if i == 3 else 0) for i, (p, g, mm, vv) in enumerate(zip(gp, gg, m, v))]
        ops.adamw_multi_(entries, hyper, ss, 1e-12)
        torch.cuda.synchronize()
        for p, r in zip(gp, ref_p):
            torch.testing.assert_close(p.cpu(), r.detach(), rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("B,d,F,k", [(128, 384, 3072, 32), (64, 384, 3072, 32), (37, 96, 400, 16),
                                     (256, 768, 6144, 32), (5, 64, 128, 8)])
def test_row_step_small_batch_form(B, d, F, k):
    """wsae_row_step (one block per row: TopK selection + decode + MSE + stamps + dv + bias gradients +
    both weight-gradient rows) against the dense K1 form's selection and a plain fp32 statement of the
    rest (model.py:114-148 and its autograd), same bf16 operand roundings as K23."""
    ops = _ops()
    torch.manual_seed(B + d)
    state = O.init_state(d, F)
    state["b_pre"] = torch.randn(d) * 0.05
    state["decoder.weight"] = state["decoder.weight"].to(torch.bfloat16).float()
    x = O.synthetic_activations(B, d, seed=B + 7)
    dev = "cuda"
    a = ops.pack_activations(x.cuda(), state["b_pre"].cuda(), 1)
    w = ops.pack_encoder(state["encoder.weight"].cuda(), state["encoder.bias"].cuda(), 1)
    idx_ref, val_ref = ops.encode_topk(a, w, B, F, d, 1, k)               # dense small-batch form of K1
    pre = ops.encode_dense(a, w, B, F, d, 1)
    w_decT = state["decoder.weight"].t().contiguous()
    stats = torch.zeros(3, dtype=torch.int64, device=dev)
    last = torch.zeros(F, dtype=torch.int64, device=dev)
    step = torch.tensor(6, dtype=torch.int64, device=dev)
    dbe, dbd, dbp = torch.zeros(F, device=dev), torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    dwe, dwd = torch.zeros(F, d, device=dev), torch.zeros(F, d, device=dev)
    resid = torch.empty(B, d, device=dev)
    dpre = torch.empty(B, k, device=dev)
    grad_out, coef = 0.5, 2.0 / (B * d)
    assert ops.row_step_supported(B, d, F, k, True)
    idx, val = ops.row_step(pre, x.cuda(), w_decT.cuda().to(torch.bfloat16), state["decoder.bias"].cuda(),
                            state["b_pre"].cuda(), torch.tensor(grad_out, device=dev), coef, k, stats=stats,
                            last_activated=last, step_count=step, d_b_enc=dbe, d_b_dec=dbd, d_w_enc=dwe,
                            d_w_decT=dwd, resid=resid, dpre_val=dpre, w_enc=state["encoder.weight"].cuda(),
                            d_b_pre=dbp)
    torch.cuda.synchronize()
    assert torch.equal(idx, idx_ref) and torch.equal(val, val_ref)
    idx_c, val_c = idx.cpu().long(), val.cpu()
    h = torch.relu(val_c)
    rows = w_decT[idx_c]                                                   # [B, k, d]
    recon = (h.to(torch.bfloat16).float().unsqueeze(-1) * rows).sum(1) + state["decoder.bias"] + state["b_pre"]
    resid_ref = recon - x
    s = coef * grad_out
    dv_ref = s * (resid_ref.to(torch.bfloat16).float().unsqueeze(1) * rows).sum(-1) * (val_c > 0)

    def close(got, want, name, rtol=2e-5, arel=2e-6):
        scale = want.abs().max().item() + 1e-30
        torch.testing.assert_close(got.cpu(), want, rtol=rtol, atol=arel * scale, msg=lambda m: f"{name}: {m}")

    close(resid, resid_ref, "resid")
    raw = stats.cpu()
    assert raw[:1].view(torch.float64).item() == pytest.approx((resid_ref.double() ** 2).sum().item(), rel=1e-6)
    assert raw[1].item() == int((val_c > 0).sum())
    want_last = torch.zeros(F, dtype=torch.int64)
    want_last[idx_c[val_c > 0]] = 7
    assert torch.equal(last.cpu(), want_last)
    close(dpre, dv_ref, "dpre", rtol=1e-4, arel=1e-5)
    close(dbe, torch.zeros(F).index_add_(0, idx_c.reshape(-1), dv_ref.reshape(-1)), "db_enc", rtol=1e-4, arel=1e-5)
    close(dbd, s * resid_ref.sum(0), "db_dec", rtol=1e-4, arel=1e-5)
    dense_dv = torch.zeros(B, F).scatter_(1, idx_c, dv_ref)
    dense_h = torch.zeros(B, F).scatter_(1, idx_c, h)
    close(dwe, dense_dv.t() @ (x - state["b_pre"]), "dW_enc", rtol=1e-4, arel=1e-5)
    close(dwd, s * dense_h.t() @ resid_ref, "dW_decT", rtol=1e-4, arel=1e-5)
    # db_pre = db_dec - db_enc . W_enc (what wsae_bpre_grad computes from the finished sums)
    close(dbp, s * resid_ref.sum(0) - dense_dv.sum(0) @ state["encoder.weight"], "db_pre", rtol=1e-4, arel=2e-5)

    # with a ticket the LAST block also does the counters kernel's job: step_count += 1, dead count against the
    # new step (model.py:183-195), {sse, l0, dead, seq} posted to the pinned host mailbox
    stats2 = torch.zeros(3, dtype=torch.int64, device=dev)
    last2 = torch.zeros(F, dtype=torch.int64, device=dev)
    last2[: F // 2] = -5                        # half the features: last fired long ago
    step2 = torch.tensor(6, dtype=torch.int64, device=dev)
    ticket = torch.zeros(1, dtype=torch.int64, device=dev)
    seq = torch.tensor([77], dtype=torch.int64, device=dev)
    mailbox = torch.zeros(4, dtype=torch.int64).pin_memory()
    ops.row_step(pre, x.cuda(), w_decT.cuda().to(torch.bfloat16), state["decoder.bias"].cuda(),
                 state["b_pre"].cuda(), torch.tensor(grad_out, device=dev), coef, k, stats=stats2,
                 last_activated=last2, step_count=step2, d_b_enc=None, d_b_dec=None, d_w_enc=None, d_w_decT=None,
                 finish=(ticket, 3, stats2[2:], seq, mailbox))
    torch.cuda.synchronize()
    assert int(step2) == 7 and int(ticket) == B
    want_last2 = torch.zeros(F, dtype=torch.int64)
    want_last2[: F // 2] = -5
    want_last2[idx_c[val_c > 0]] = 7
    dead = int(((7 - want_last2) > 3).sum())
    assert torch.equal(last2.cpu(), want_last2)
    assert int(stats2[2]) == dead
    assert mailbox[3].item() == 77 and mailbox[2].item() == dead and mailbox[1].item() == int((val_c > 0).sum())
    assert mailbox[:1].view(torch.float64).item() == pytest.approx((resid_ref.double() ** 2).sum().item(), rel=1e-6)
