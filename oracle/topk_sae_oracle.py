"""CPU oracle for the TopK-SAE train step — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  The product path (``whisper_sae_b200``) never does and has no CPU
fallback.

This is a functional restatement (plain tensors, explicit forward AND explicit backward, explicit
clip / AdamW / renorm / counters — no ``nn.Module``, no autograd) of what the reference computes:

    /root/reference/src/whisper_sae/sae/model.py      TopKSAE  (:38-89 init, :91-96 renorm,
                                                      :98-118 encode, :120-129 decode,
                                                      :131-166 forward, :168-195 dead features,
                                                      :197-257 resampling)
    /root/reference/src/whisper_sae/sae/training.py   SAETrainer.train_step (:161-217),
                                                      setup_scheduler (:136-159)

The reference's arithmetic lives in third-party torch (pyproject pins torch>=2.1.0; installed here:
torch 2.11.0), so the oracle uses the same CPU torch ops for the dense contractions and restates
the *algorithm*; its backward is the hand-derived gradient the reference obtains from autograd.

Parity pinning: ``oracle/make_golden.py`` imports the live reference from /root/reference/src in
the build container and writes golden vectors to ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this oracle against them (the reference's own tests hold no golden vectors — SURVEY §8c).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
from torch import Tensor

PARAM_ORDER = ("b_pre", "encoder.weight", "encoder.bias", "decoder.weight", "decoder.bias")


# --------------------------------------------------------------------------------------------
# construction (model.py:38-89)
# --------------------------------------------------------------------------------------------
def init_state(input_dim: int, hidden_dim: int, dtype=torch.float32) -> dict[str, Tensor]:
    """Fresh TopKSAE state_dict drawn from the *current* torch RNG exactly as the reference
    constructor does: nn.Linear(d,F) then nn.Linear(F,d) default inits (kaiming-uniform a=sqrt(5)
    weights, uniform(+-1/sqrt(fan_in)) biases), xavier_uniform_ on the decoder, unit-norm columns,
    x0.1 (model.py:63-64,81-89)."""

    def linear_init(out_f: int, in_f: int) -> tuple[Tensor, Tensor]:
        w = torch.empty(out_f, in_f)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(in_f)
        b = torch.empty(out_f).uniform_(-bound, bound)
        return w, b

    w_enc, b_enc = linear_init(hidden_dim, input_dim)
    w_dec, b_dec = linear_init(input_dim, hidden_dim)
    torch.nn.init.xavier_uniform_(w_dec)
    w_dec = torch.nn.functional.normalize(w_dec, dim=0) * 0.1
    return {
        "b_pre": torch.zeros(input_dim, dtype=dtype),
        "feature_last_activated": torch.zeros(hidden_dim, dtype=torch.long),
        "step_count": torch.tensor(0, dtype=torch.long),
        "encoder.weight": w_enc.to(dtype),
        "encoder.bias": b_enc.to(dtype),
        "decoder.weight": w_dec.to(dtype),
        "decoder.bias": b_dec.to(dtype),
    }


def _bf16_round(t: Tensor) -> Tensor:
    return t.to(torch.bfloat16).to(t.dtype)


# --------------------------------------------------------------------------------------------
# forward (model.py:98-181)
# --------------------------------------------------------------------------------------------
@dataclass
class Forward:
    idx: Tensor          # [B,k] selected feature indices (torch.topk order)
    val: Tensor          # [B,k] signed pre-activations at idx
    pre: Tensor          # [B,F] dense pre-activations (oracle only; the kernels never build it)
    recon: Tensor        # [B,d]
    resid: Tensor        # recon - target
    loss: Tensor         # mean((recon - target)^2)
    l0: Tensor           # mean_b #(hidden > 0)
    w_dec_used: Tensor   # decoder weights as seen by decode/backward ([d,F]; bf16-rounded in bf16 mode)
    xc_used: Tensor      # centred input as seen by the encoder GEMM


def pre_activations(state: dict[str, Tensor], x: Tensor, quantize: str | None = None) -> tuple[Tensor, Tensor]:
    """(x - b_pre) @ W_enc^T + b_enc (model.py:108,111).  quantize="bf16" rounds both GEMM operands
    to bf16 (fp32 accumulate, fp32 bias) the way the tensor-core path sees them."""
    xc = x - state["b_pre"]
    w = state["encoder.weight"]
    if quantize == "bf16":
        xc_q, w_q = _bf16_round(xc), _bf16_round(w)
    else:
        xc_q, w_q = xc, w
    return xc_q @ w_q.t() + state["encoder.bias"], xc_q


def forward(state: dict[str, Tensor], x: Tensor, k: int, *, target: Tensor | None = None,
            training: bool = True, quantize: str | None = None, rows_total: int | None = None) -> Forward:
    pre, xc_q = pre_activations(state, x, quantize)
    val, idx = torch.topk(pre, k, dim=-1)                               # model.py:114
    h = torch.relu(val)                                                 # model.py:116
    w_dec = state["decoder.weight"]                                     # [d, F]
    w_used = _bf16_round(w_dec) if quantize == "bf16" else w_dec
    rows = w_used.t()[idx]                                              # [B,k,d] gather == hidden @ W_dec^T
    recon = (h.unsqueeze(-1) * rows).sum(dim=1) + state["decoder.bias"]
    if "b_pre" in state and state["b_pre"] is not None:
        recon = recon + state["b_pre"]                                  # model.py:129
    tgt = x if target is None else target
    resid = recon - tgt
    n_rows = rows_total if rows_total is not None else x.shape[0]
    loss = (resid.double() ** 2).sum().to(resid.dtype) / (n_rows * tgt.shape[1])   # model.py:145
    l0 = (h > 0).to(resid.dtype).sum(dim=-1).sum() / n_rows             # model.py:148
    if training:                                                        # model.py:168-181
        state["step_count"] += 1
        fired = idx[h > 0]
        state["feature_last_activated"][fired] = state["step_count"]
    return Forward(idx, val, pre, recon, resid, loss, l0, w_used, xc_q)


def dense_hidden(fwd: Forward, hidden_dim: int) -> Tensor:
    hidden = torch.zeros(fwd.idx.shape[0], hidden_dim, dtype=fwd.val.dtype)
    hidden.scatter_(-1, fwd.idx, torch.relu(fwd.val))                   # model.py:115-116
    return hidden


def dead_features(state: dict[str, Tensor], threshold: int) -> Tensor:
    return (state["step_count"] - state["feature_last_activated"]) > threshold   # model.py:183-191


def dead_feature_ratio(state: dict[str, Tensor], threshold: int) -> float:
    return dead_features(state, threshold).float().mean().item()                 # model.py:193-195


def resample_dead_features(state: dict[str, Tensor], inputs: Tensor, k: int, threshold: int,
                           num_resample: int | None = None, training: bool = True) -> int:
    """TopKSAE.resample_dead_features (model.py:197-257), in place on ``state``.

    Order of events as in the reference: the dead set is taken BEFORE the no-grad forward; that
    forward bumps ``step_count`` and stamps the fired features when the module is in train mode
    (model.py:229 -> :168-181); the i-th dead feature (ascending index) then receives the i-th
    highest-error input row, L2-normalised (F.normalize, eps 1e-12), as encoder row AND decoder
    column, bias 0, ``feature_last_activated = step_count``.  Returns ``num_dead`` (capped by
    ``num_resample``) even when there are fewer input rows than dead features (:239-257)."""
    dead_idx = torch.where(dead_features(state, threshold))[0]
    num_dead = int(dead_idx.numel())
    if num_dead == 0:
        return 0
    if num_resample is not None:
        num_dead = min(num_dead, num_resample)
        dead_idx = dead_idx[:num_dead]
    fwd = forward(state, inputs, k, training=training)
    errors = ((inputs - fwd.recon) ** 2).sum(dim=-1)
    n = min(num_dead, errors.numel())
    _, top = torch.topk(errors, n)
    rows = torch.nn.functional.normalize(inputs[top], dim=-1)
    tgt = dead_idx[:n]
    state["encoder.weight"][tgt] = rows
    state["encoder.bias"][tgt] = 0.0
    w_dec = state["decoder.weight"]
    w_dec[:, tgt] = rows.t()
    state["feature_last_activated"][tgt] = state["step_count"]
    return num_dead


# --------------------------------------------------------------------------------------------
# backward: what autograd derives for loss = mean((recon - x)^2)  (run at training.py:184)
# --------------------------------------------------------------------------------------------
def backward(state: dict[str, Tensor], x: Tensor, fwd: Forward, grad_out: float = 1.0,
             rows_total: int | None = None, same_target: bool = True) -> dict[str, Tensor]:
    B, d_out = fwd.resid.shape
    n_rows = rows_total if rows_total is not None else B
    F = state["encoder.weight"].shape[0]
    g = fwd.resid * (2.0 * grad_out / (n_rows * d_out))                 # dL/drecon
    h = torch.relu(fwd.val)
    rows = fwd.w_dec_used.t()[fwd.idx]                                  # [B,k,d]
    dv = (rows * g.unsqueeze(1)).sum(-1) * (fwd.val > 0).to(g.dtype)    # [B,k]
    flat_idx = fwd.idx.reshape(-1)
    d_w_decT = torch.zeros(F, d_out, dtype=g.dtype)
    d_w_decT.index_add_(0, flat_idx, (h.unsqueeze(-1) * g.unsqueeze(1)).reshape(-1, d_out))
    xc = x - state["b_pre"]
    d_w_enc = torch.zeros_like(state["encoder.weight"])
    d_w_enc.index_add_(0, flat_idx, (dv.unsqueeze(-1) * xc.unsqueeze(1)).reshape(-1, xc.shape[1]))
    d_b_enc = torch.zeros(F, dtype=g.dtype).index_add_(0, flat_idx, dv.reshape(-1))
    d_b_dec = g.sum(0)
    d_b_pre = -(d_b_enc @ state["encoder.weight"])
    if same_target:
        d_b_pre = d_b_pre + d_b_dec
    dpre = torch.zeros(B, F, dtype=g.dtype).scatter_(-1, fwd.idx, dv)
    dx = dpre @ state["encoder.weight"] - (g if same_target else 0.0)
    return {"b_pre": d_b_pre, "encoder.weight": d_w_enc, "encoder.bias": d_b_enc,
            "decoder.weight": d_w_decT.t().contiguous(), "decoder.bias": d_b_dec,
            "dx": dx, "dpre_val": dv}


# --------------------------------------------------------------------------------------------
# optimiser side of train_step (training.py:187-202)
# --------------------------------------------------------------------------------------------
def clip_grad_norm_(grads: dict[str, Tensor], max_norm: float) -> Tensor:
    """torch.nn.utils.clip_grad_norm_ (training.py:188-191): scale by min(1, max_norm/(norm+1e-6))."""
    total = torch.sqrt(sum((grads[n].double() ** 2).sum() for n in PARAM_ORDER)).to(torch.float32)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for n in PARAM_ORDER:
        grads[n] = grads[n] * coef.to(grads[n].dtype)
    return total


@dataclass
class AdamWState:
    step: int = 0
    exp_avg: dict[str, Tensor] = field(default_factory=dict)
    exp_avg_sq: dict[str, Tensor] = field(default_factory=dict)


def adamw_step(state: dict[str, Tensor], grads: dict[str, Tensor], opt: AdamWState, lr: float,
               betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0) -> None:
    """torch.optim.AdamW single-tensor update (training.py:63-67,193)."""
    opt.step += 1
    b1, b2 = betas
    bc1 = 1.0 - b1 ** opt.step
    bc2_sqrt = math.sqrt(1.0 - b2 ** opt.step)
    for n in PARAM_ORDER:
        p, g = state[n], grads[n]
        if n not in opt.exp_avg:
            opt.exp_avg[n] = torch.zeros_like(p)
            opt.exp_avg_sq[n] = torch.zeros_like(p)
        m, v = opt.exp_avg[n], opt.exp_avg_sq[n]
        p.mul_(1.0 - lr * weight_decay)
        m.lerp_(g, 1.0 - b1)
        v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-(lr / bc1))


def renorm_decoder_(state: dict[str, Tensor]) -> None:
    """F.normalize(W_dec, dim=0): columns / max(||col||, 1e-12) (model.py:91-96)."""
    w = state["decoder.weight"]
    state["decoder.weight"] = w / w.norm(dim=0, keepdim=True).clamp_min(1e-12)


def lr_sequence(base_lr: float, total_steps: int, warmup_steps: int) -> list[float]:
    """lr used at step 1..total_steps under setup_scheduler (training.py:136-159):
    SequentialLR(LinearLR(0.01->1, warm), CosineAnnealingLR(T_max=total-warm, eta_min=0.1*lr)),
    warm = min(warmup_steps, total_steps // 10).  Restated in closed form."""
    warm = min(warmup_steps, total_steps // 10)
    eta_min = 0.1 * base_lr
    t_max = total_steps - warm
    out = []
    for s in range(total_steps):
        if s < warm:
            out.append(base_lr * (0.01 + (1.0 - 0.01) * s / warm))
        else:
            t = s - warm
            out.append(eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * t / t_max)) / 2)
    return out


@dataclass
class StepResult:
    loss: float
    l0: float
    dead_feature_ratio: float
    grad_norm: float
    fwd: Forward
    grads: dict[str, Tensor]


def train_step(state: dict[str, Tensor], opt: AdamWState, x: Tensor, k: int, lr: float, *,
               gradient_clip: float = 1.0, weight_decay: float = 0.0, dead_threshold: int = 10_000,
               quantize: str | None = None) -> StepResult:
    """One SAETrainer.train_step (training.py:161-217), fp32, AMP off (as the trainer forces on CPU)."""
    fwd = forward(state, x, k, training=True, quantize=quantize)
    grads = backward(state, x, fwd)
    raw = {n: grads[n].clone() for n in PARAM_ORDER}
    norm = clip_grad_norm_(grads, gradient_clip)
    adamw_step(state, grads, opt, lr, weight_decay=weight_decay)
    renorm_decoder_(state)
    return StepResult(fwd.loss.item(), fwd.l0.item(), dead_feature_ratio(state, dead_threshold),
                      norm.item(), fwd, raw)


# --------------------------------------------------------------------------------------------
# synthetic activations: row-standardised Gaussians == final-LayerNorm output of a
# random-init Whisper (gamma=1, beta=0)
# --------------------------------------------------------------------------------------------
def synthetic_activations(n_rows: int, d: int, seed: int) -> Tensor:
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_rows, d, generator=g)
    return (x - x.mean(1, keepdim=True)) / x.std(1, unbiased=False, keepdim=True)


def near_tie_rows(pre: Tensor, k: int, tau: float) -> Tensor:
    """Rows whose k-th and (k+1)-th largest pre-activations are within tau (documented near-ties)."""
    if pre.shape[1] <= k:
        return torch.zeros(pre.shape[0], dtype=torch.bool)
    top = torch.topk(pre, k + 1, dim=-1).values
    return (top[:, k - 1] - top[:, k]).abs() <= tau
