"""K0 + K1 parity: packed operands and the tcgen05 GEMM with fused TopK, through the C ABI,
against the CPU oracle (model.py:108-114 semantics)."""

import pytest
import torch

from oracle import topk_sae_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from whisper_sae_b200 import ops
    return ops


def _split_piece(v: torch.Tensor, p: int) -> torch.Tensor:
    h0 = v.to(torch.bfloat16)
    if p == 0:
        return h0
    r1 = v - h0.float()
    h1 = r1.to(torch.bfloat16)
    if p == 1:
        return h1
    return (r1 - h1.float()).to(torch.bfloat16)


@pytest.mark.parametrize("terms", [1, 3, 6])
def test_pack_layout(terms):
    ops = _ops()
    torch.manual_seed(1)
    B, d = 37, 40
    x = torch.randn(B, d, device="cuda")
    b_pre = torch.randn(d, device="cuda") * 0.1
    ps = ops.packed_shape(d, terms)
    assert ps.dp == 40 and ps.used_cols == terms * 40 + 16 and ps.kp % 64 == 0
    a = ops.pack_activations(x, b_pre, terms)
    assert a.shape == (128, ps.kp)
    sched = {1: [0], 3: [0, 1, 0], 6: [0, 0, 1, 0, 1, 2]}[terms]
    xc = x - b_pre
    for t, p in enumerate(sched):
        assert torch.equal(a[:B, t * ps.dp: t * ps.dp + d], _split_piece(xc, p))
    bias_blk = a[:B, terms * ps.dp: terms * ps.dp + 16].float()
    assert torch.equal(bias_blk[:, :3], torch.ones(B, 3, device="cuda"))
    assert (bias_blk[:, 3:] == 0).all() and (a[B:] == 0).all() and (a[:, ps.used_cols:] == 0).all()

    F = 50
    w = torch.randn(F, d, device="cuda")
    b = torch.randn(F, device="cuda")
    wp = ops.pack_encoder(w, b, terms)
    assert wp.shape == (256, ps.kp)
    wsched = {1: [0], 3: [0, 0, 1], 6: [0, 1, 0, 2, 1, 0]}[terms]
    for t, p in enumerate(wsched):
        assert torch.equal(wp[:F, t * ps.dp: t * ps.dp + d], _split_piece(w, p))
    bb = wp[:F, terms * ps.dp: terms * ps.dp + 3].float().sum(-1)
    torch.testing.assert_close(bb, b, rtol=1e-6, atol=1e-7)


def _encode(x, state, k, terms, nsplit=None):
    ops = _ops()
    dev = "cuda"
    xg = x.to(dev)
    a = ops.pack_activations(xg, state["b_pre"].to(dev), terms)
    w = ops.pack_encoder(state["encoder.weight"].to(dev), state["encoder.bias"].to(dev), terms)
    F, d = state["encoder.weight"].shape
    idx, val = ops.encode_topk(a, w, x.shape[0], F, d, terms, k, nsplit=nsplit)
    torch.cuda.synchronize()
    return idx.cpu(), val.cpu()


def _compare_sets(idx, val, pre, k, tau, val_rtol, val_atol):
    """Index sets equal to torch.topk(pre) except rows with a documented near-tie at the k-th place."""
    B, F = pre.shape
    ref_val, ref_idx = torch.topk(pre, k, dim=-1)
    got = torch.sort(idx.long(), -1).values
    want = torch.sort(ref_idx, -1).values
    assert (idx >= 0).all() and (idx < F).all()
    assert all(len(set(r.tolist())) == k for r in idx), "duplicate indices in a row"
    bad = (got != want).any(-1)
    ties = O.near_tie_rows(pre, k, tau)
    unexplained = bad & ~ties
    assert not unexplained.any(), (
        f"{int(unexplained.sum())}/{B} rows differ from torch.topk beyond the near-tie rule "
        f"(rows with near ties: {int(ties.sum())}); first bad row {int(unexplained.nonzero()[0])}")
    # values must be the pre-activations at the returned indices
    torch.testing.assert_close(val, pre.gather(-1, idx.long()), rtol=val_rtol, atol=val_atol)
    return int(bad.sum())


SHAPES = [
    # B, d, F, k
    (1, 32, 128, 4),
    (10, 64, 256, 32),
    (32, 64, 256, 8),
    (10, 32, 32, 32),      # k == F (identity test shape, tests/test_sae_model.py:515-536)
    (64, 384, 3072, 32),   # BASELINE config 1 (tiny_test.yaml)
    (300, 128, 1024, 32),
    (130, 96, 400, 16),    # ragged: F not a multiple of 256, B not a multiple of 128
    (16, 64, 128, 64),     # k > 32 variant of the kernel
]


@pytest.mark.parametrize("B,d,F,k", SHAPES)
def test_topk_fp32_grade(B, d, F, k):
    torch.manual_seed(B * 7 + F)
    state = O.init_state(d, F)
    state["b_pre"] = torch.randn(d) * 0.05
    x = O.synthetic_activations(B, d, seed=B + d)
    pre, _ = O.pre_activations(state, x)
    idx, val = _encode(x, state, k, terms=6)
    tau = 1e-5 * pre.abs().max().item()
    _compare_sets(idx, val, pre, k, tau, val_rtol=1e-5, val_atol=2e-6)


@pytest.mark.parametrize("B,d,F,k", SHAPES)
def test_topk_bf16_matches_bf16_oracle(B, d, F, k):
    """bf16 mode == oracle with bf16-rounded GEMM operands (fp32 accumulate, exact bias)."""
    torch.manual_seed(B * 11 + F)
    state = O.init_state(d, F)
    x = O.synthetic_activations(B, d, seed=B + d + 1)
    pre_q, _ = O.pre_activations(state, x, quantize="bf16")
    idx, val = _encode(x, state, k, terms=1)
    tau = 1e-5 * pre_q.abs().max().item()
    _compare_sets(idx, val, pre_q, k, tau, val_rtol=1e-5, val_atol=2e-6)
    # and against the unquantised oracle only the documented bf16 near-tie rule applies
    pre, _ = O.pre_activations(state, x)
    _compare_sets(idx, val, pre, k, tau=2.0 ** -6 * pre.abs().max().item(), val_rtol=0.05, val_atol=0.05)


@pytest.mark.parametrize("nsplit", [1, 2, 3, 12])
def test_f_split_merge(nsplit):
    torch.manual_seed(3)
    B, d, F, k = 200, 64, 3072, 32
    state = O.init_state(d, F)
    x = O.synthetic_activations(B, d, seed=5)
    pre, _ = O.pre_activations(state, x)
    idx, val = _encode(x, state, k, terms=6, nsplit=nsplit)
    _compare_sets(idx, val, pre, k, 1e-5 * pre.abs().max().item(), 1e-5, 2e-6)


def test_ties_resolve_to_lowest_index():
    """All pre-activations equal (x = 0, constant bias): any k indices are a valid torch.topk
    answer; ours are deterministic — the k lowest."""
    d, F, k, B = 64, 512, 8, 5
    state = O.init_state(d, F)
    state["encoder.bias"] = torch.full((F,), 0.25)
    x = torch.zeros(B, d)
    for ns in (1, 2):
        idx, val = _encode(x, state, k, terms=1, nsplit=ns)
        assert torch.equal(torch.sort(idx, -1).values, torch.arange(k, dtype=torch.int32).expand(B, k))
        assert (val == 0.25).all()


def test_large_v3_shape_and_many_rows():
    """BASELINE config 4 geometry (1280 -> 40960) on a few hundred rows, and a many-row tiny case."""
    torch.manual_seed(0)
    d, F, k, B = 1280, 40960, 32, 256
    state = {"b_pre": torch.zeros(d), "encoder.weight": torch.randn(F, d) / d ** 0.5,
             "encoder.bias": torch.randn(F) * 0.01}
    x = O.synthetic_activations(B, d, seed=9)
    pre_q, _ = O.pre_activations(state, x, quantize="bf16")
    idx, val = _encode(x, state, k, terms=1)
    _compare_sets(idx, val, pre_q, k, 2e-5 * pre_q.abs().max().item(), 2e-5, 1e-5)

    d, F, B = 384, 3072, 20000
    state = O.init_state(d, F)
    x = O.synthetic_activations(B, d, seed=10)
    pre_q, _ = O.pre_activations(state, x, quantize="bf16")
    idx, val = _encode(x, state, k, terms=1)
    _compare_sets(idx, val, pre_q, k, 1e-5 * pre_q.abs().max().item(), 1e-5, 2e-6)


def test_rejects_cpu_and_bad_k():
    ops = _ops()
    with pytest.raises(RuntimeError):
        ops.pack_activations(torch.randn(4, 32), None, 1)
    a = ops.pack_activations(torch.randn(4, 32, device="cuda"), None, 1)
    w = ops.pack_encoder(torch.randn(16, 32, device="cuda"), None, 1)
    with pytest.raises(RuntimeError):
        ops.encode_topk(a, w, 4, 16, 32, 1, k=17)


@pytest.mark.parametrize("B,d,F,k", [(128, 384, 3072, 32), (64, 384, 3072, 32), (130, 96, 400, 16),
                                     (1000, 128, 1024, 32), (16, 64, 128, 64), (256, 768, 6144, 32)])
def test_small_batch_dense_form_equals_fused(B, d, F, k):
    """wsae_encode_topk_dense (GEMM -> [B,F] scratch -> radix select per row, the route of batches up
    to 1024 rows) returns the same index sets and bit-identical values as the fused epilogue."""
    ops = _ops()
    torch.manual_seed(B + F)
    state = O.init_state(d, F)
    state["b_pre"] = torch.randn(d) * 0.05
    x = O.synthetic_activations(B, d, seed=B + 3)
    for terms in (1, 6):
        a = ops.pack_activations(x.cuda(), state["b_pre"].cuda(), terms)
        w = ops.pack_encoder(state["encoder.weight"].cuda(), state["encoder.bias"].cuda(), terms)
        i_d, v_d = ops.encode_topk(a, w, B, F, d, terms, k)            # dense route (B <= 1024)
        i_f, v_f = ops.encode_topk(a, w, B, F, d, terms, k, nsplit=1)  # fused epilogue
        torch.cuda.synchronize()
        o_d, o_f = torch.sort(i_d, -1), torch.sort(i_f, -1)
        assert torch.equal(o_d.values, o_f.values)
        assert torch.equal(v_d.gather(-1, o_d.indices), v_f.gather(-1, o_f.indices))
        assert torch.equal(o_d.indices, torch.arange(k, device="cuda").expand(B, k)), "ascending feature order"


def test_small_batch_dense_form_nan_rows_and_ties():
    """Rows with NaN inputs: NaN pre-activations are never selected, the list is padded with
    (-inf, -1) like the fused epilogue; equal values keep the lowest indices."""
    ops = _ops()
    d, F, k, B = 64, 512, 8, 6
    state = O.init_state(d, F)
    state["encoder.bias"] = torch.full((F,), 0.25)
    x = torch.zeros(B, d)
    x[2] = float("nan")
    a = ops.pack_activations(x.cuda(), None, 1)
    w = ops.pack_encoder(state["encoder.weight"].cuda(), state["encoder.bias"].cuda(), 1)
    idx, val = ops.encode_topk(a, w, B, F, d, 1, k)
    idx_f, val_f = ops.encode_topk(a, w, B, F, d, 1, k, nsplit=1)
    torch.cuda.synchronize()
    good = [0, 1, 3, 4, 5]
    assert torch.equal(idx[good].cpu(), torch.arange(k, dtype=torch.int32).expand(len(good), k))
    assert (val[good] == 0.25).all()
    assert (idx[2] == -1).all() and torch.isinf(val[2]).all() and (val[2] < 0).all()
    assert torch.equal(torch.sort(idx_f, -1).values, torch.sort(idx, -1).values)
