"""Generate golden vectors from the LIVE reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Imports ``whisper_sae`` from /root/reference/src (read-only, never copied), drives the reference's
own ``TopKSAE`` + ``SAETrainer`` (CPU, fp32, AMP off — training.py:73-75) on seeded synthetic
activations and stores what they produced under ``tests/golden/``.  The GPU box has no
/root/reference: tests there read only these fixtures.

Each fixture is a dict: the recipe (seeds, shapes, hyper-parameters) + the reference's outputs.
Inputs are regenerated from the seeds with ``oracle.topk_sae_oracle.synthetic_activations``.
"""

from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import torch

REF_SRC = Path("/root/reference/src")
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF_SRC))
sys.dont_write_bytecode = True

from oracle.topk_sae_oracle import PARAM_ORDER, synthetic_activations  # noqa: E402
from whisper_sae.config import TrainingConfig  # noqa: E402  (reference)
from whisper_sae.sae.model import TopKSAE  # noqa: E402  (reference)
from whisper_sae.sae.training import SAETrainer  # noqa: E402  (reference)

GOLDEN = ROOT / "tests" / "golden"


def digest(t: torch.Tensor) -> dict:
    td = t.detach().double().reshape(-1)
    stride = max(1, td.numel() // 2048)
    return {
        "shape": tuple(t.shape),
        "sum": td.sum().item(),
        "abs_sum": td.abs().sum().item(),
        "sumsq": (td * td).sum().item(),
        "sample_stride": stride,
        "sample": t.detach().reshape(-1)[::stride].clone(),
    }


def run_case(name: str, d: int, F: int, k: int, B: int, steps: int, total_steps: int, lr: float,
             warmup: int, thr: int, model_seed: int, data_seed: int, full_state: bool) -> None:
    torch.manual_seed(model_seed)
    model = TopKSAE(d, F, k=k, dead_feature_threshold=thr)
    init_state = {n: v.clone() for n, v in model.state_dict().items()}
    cfg = TrainingConfig(batch_size=B, learning_rate=lr, warmup_steps=warmup, epochs=1,
                         use_amp=False, num_workers=0)
    with tempfile.TemporaryDirectory() as tmp:
        trainer = SAETrainer(model, cfg, device="cpu", run_dir=Path(tmp))
        trainer.setup_scheduler(total_steps)
        x_all = synthetic_activations(B * steps, d, data_seed)
        per_step = []
        first = {}
        for s in range(steps):
            xb = x_all[s * B:(s + 1) * B]
            lr_used = trainer.optimizer.param_groups[0]["lr"]
            if s == 0:
                # step-1 internals, recomputed through the reference module on a clone
                probe = TopKSAE(d, F, k=k, dead_feature_threshold=thr)
                probe.load_state_dict(model.state_dict())
                probe.train()
                out = probe(xb)
                out.loss.backward()
                pre = probe.encoder(xb - probe.b_pre)
                tv, ti = torch.topk(pre, k, dim=-1)
                first = {
                    "topk_idx_sorted": torch.sort(ti, dim=-1).values.to(torch.int32),
                    "kth_gap": (torch.topk(pre, min(k + 1, F), dim=-1).values[:, k - 1]
                                - (torch.topk(pre, min(k + 1, F), dim=-1).values[:, -1])).detach(),
                    "pre_absmax": pre.abs().max().item(),
                    "loss": out.loss.item(),
                    "l0": out.l0.item(),
                    "grads": {n: (dict(probe.named_parameters())[n].grad.clone() if full_state
                                  else digest(dict(probe.named_parameters())[n].grad))
                              for n in PARAM_ORDER},
                }
            with torch.no_grad():   # k-th / (k+1)-th gap of this step's pre-activations (near-tie rule)
                pre_s = model.encoder(xb - model.b_pre)
                top = torch.topk(pre_s, min(k + 1, F), dim=-1).values
                min_gap_rel = ((top[:, k - 1] - top[:, -1]).min() / pre_s.abs().max()).item()
            m = trainer.train_step(xb)
            per_step.append({"loss": m.loss, "l0": m.l0, "dead_feature_ratio": m.dead_feature_ratio,
                             "lr_used": lr_used, "lr_reported": m.learning_rate, "step": m.step,
                             "min_gap_rel": min_gap_rel})
        final = model.state_dict()
        fixture = {
            "recipe": dict(name=name, d=d, F=F, k=k, B=B, steps=steps, total_steps=total_steps, lr=lr,
                           warmup=warmup, dead_threshold=thr, model_seed=model_seed,
                           data_seed=data_seed, gradient_clip=cfg.gradient_clip,
                           weight_decay=cfg.weight_decay),
            "torch_version": torch.__version__,
            "threads": torch.get_num_threads(),
            "init_digest": {n: digest(v) for n, v in init_state.items() if v.is_floating_point()},
            "first_step": first,
            "per_step": per_step,
            "final_counters": {"feature_last_activated": final["feature_last_activated"].clone(),
                               "step_count": final["step_count"].clone()},
            "final_params": {n: (final[n].clone() if full_state else digest(final[n]))
                             for n in PARAM_ORDER},
        }
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.save(fixture, GOLDEN / f"{name}.pt")
    print(f"{name}: losses {[round(p['loss'], 6) for p in per_step]}")


def run_dead_case() -> None:
    """tests/test_sae_model.py:251-294 scenario: one fixed row, k=4 of 128, threshold 50, 60 steps."""
    torch.manual_seed(7)
    model = TopKSAE(32, 128, k=4, dead_feature_threshold=50)
    model.train()
    g = torch.Generator().manual_seed(99)
    x = torch.randn(1, 32, generator=g)
    with torch.no_grad():
        for _ in range(60):
            model(x)
    dead = model.get_dead_features()
    fixture = {
        "recipe": dict(name="dead_fixed_row", d=32, F=128, k=4, thr=50, steps=60, model_seed=7, x_seed=99),
        "num_alive": int((~dead).sum()),
        "dead_mask": dead.clone(),
        "feature_last_activated": model.feature_last_activated.clone(),
        "step_count": int(model.step_count),
    }
    torch.save(fixture, GOLDEN / "dead_fixed_row.pt")
    print("dead_fixed_row: alive", fixture["num_alive"])


def resample_pre_state(d: int, F: int, k: int, thr: int, model_seed: int, counter_seed: int,
                       step_count: int) -> dict:
    """State before a resample call, reproducible from seeds alone (the GPU tests rebuild it the same
    way): seeded module init, then a seeded pattern of ``feature_last_activated`` stamps in
    [0, step_count) so that roughly (1 - thr / step_count) of the features count as dead."""
    torch.manual_seed(model_seed)
    model = TopKSAE(d, F, k=k, dead_feature_threshold=thr)
    g = torch.Generator().manual_seed(counter_seed)
    with torch.no_grad():
        model.feature_last_activated.copy_(torch.randint(0, step_count, (F,), generator=g))
        model.step_count.fill_(step_count)
    return model


def run_resample_case(name: str, d: int, F: int, k: int, thr: int, rows: int, num_resample,
                      model_seed: int, counter_seed: int, data_seed: int, step_count: int,
                      train_mode: bool = True) -> None:
    """model.py:197-257 driven on the live reference module (train mode: the forward inside bumps
    step_count, :229)."""
    model = resample_pre_state(d, F, k, thr, model_seed, counter_seed, step_count)
    model.train(train_mode)
    dead_before = torch.where(model.get_dead_features())[0]
    x = synthetic_activations(rows, d, data_seed)
    ret = model.resample_dead_features(x, num_resample=num_resample)
    n_cap = len(dead_before) if num_resample is None else min(len(dead_before), num_resample)
    n_written = min(n_cap, rows)
    tgt = dead_before[:n_written]
    sd = model.state_dict()
    fixture = {
        "recipe": dict(name=name, d=d, F=F, k=k, thr=thr, rows=rows, num_resample=num_resample,
                       model_seed=model_seed, counter_seed=counter_seed, data_seed=data_seed,
                       step_count=step_count, train_mode=train_mode),
        "torch_version": torch.__version__,
        "returned": int(ret),
        "dead_before": dead_before.clone(),
        "written": tgt.clone(),
        "encoder_rows": sd["encoder.weight"][tgt].clone(),          # the rewritten rows, in full
        "decoder_cols_T": sd["decoder.weight"][:, tgt].t().contiguous().clone(),
        "encoder_bias_written": sd["encoder.bias"][tgt].clone(),
        "feature_last_activated": sd["feature_last_activated"].clone(),
        "step_count": int(sd["step_count"]),
        "after_digest": {n: digest(sd[n]) for n in PARAM_ORDER},
    }
    torch.save(fixture, GOLDEN / f"{name}.pt")
    print(f"{name}: returned {ret}, dead before {len(dead_before)}, rows written {n_written}")


if __name__ == "__main__":
    torch.set_num_threads(8)
    run_case("small_64x256", d=64, F=256, k=8, B=32, steps=6, total_steps=40, lr=1e-3, warmup=4,
             thr=3, model_seed=42, data_seed=1234, full_state=True)
    run_case("tiny_test_384x3072", d=384, F=3072, k=32, B=64, steps=4, total_steps=1000, lr=1e-4,
             warmup=100, thr=1000, model_seed=42, data_seed=1234, full_state=False)
    run_case("mid_128x1024_k32", d=128, F=1024, k=32, B=256, steps=3, total_steps=100, lr=3e-4,
             warmup=10, thr=2, model_seed=5, data_seed=77, full_state=False)
    run_dead_case()
    # BASELINE configs 3 / 4 (config.py:25-33 whisper-small d=768; :45-50 expansion_factor le=32)
    run_case("small_768x6144", d=768, F=6144, k=32, B=256, steps=3, total_steps=1000, lr=1e-4,
             warmup=100, thr=1000, model_seed=42, data_seed=2345, full_state=False)
    run_case("large_1280x40960", d=1280, F=40960, k=32, B=256, steps=2, total_steps=1000, lr=1e-4,
             warmup=100, thr=1000, model_seed=42, data_seed=3456, full_state=False)
    # resample_dead_features (model.py:197-257)
    run_resample_case("resample_64x256", d=64, F=256, k=8, thr=10, rows=16, num_resample=None,
                      model_seed=3, counter_seed=4, data_seed=5, step_count=40)
    run_resample_case("resample_64x256_eval", d=64, F=256, k=8, thr=10, rows=128, num_resample=20,
                      model_seed=3, counter_seed=4, data_seed=5, step_count=40, train_mode=False)
    run_resample_case("resample_384x3072", d=384, F=3072, k=32, thr=50, rows=1024, num_resample=64,
                      model_seed=42, counter_seed=9, data_seed=4567, step_count=100)
    run_resample_case("resample_1280x40960", d=1280, F=40960, k=32, thr=50, rows=2048, num_resample=32,
                      model_seed=42, counter_seed=9, data_seed=5678, step_count=100)
