"""Worker of tests/test_parallel.py::test_nccl_two_rank_step_*: one process per GPU (torchrun),
real NCCL process group.  Every rank trains the batch-sharded step (`data_parallel=True`) on its
row shard and checks it against (a) the live-reference golden trace in fp32-grade mode and (b) the
single-device step on the concatenated batch in bf16 mode.  Writes `<out>/rank<r>.json`."""

import json
import os
import sys
import tempfile
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import topk_sae_oracle as O  # noqa: E402  (checker only)
from tests.conftest import load_golden  # noqa: E402
from whisper_sae_b200.config import TrainingConfig  # noqa: E402
from whisper_sae_b200.sae import SAETrainer, TopKSAE, parallel  # noqa: E402


def _trainer(r, use_amp, dev, **kw):
    torch.manual_seed(r["model_seed"])
    sae = TopKSAE(r["d"], r["F"], k=r["k"], dead_feature_threshold=r["dead_threshold"])
    cfg = TrainingConfig(batch_size=r["B"], learning_rate=r["lr"], warmup_steps=r["warmup"], epochs=1,
                         use_amp=use_amp, num_workers=0)
    tr = SAETrainer(sae, cfg, device=dev, run_dir=Path(tempfile.mkdtemp()), **kw)
    tr.setup_scheduler(r["total_steps"])
    return tr


def golden_fp32(name: str, rank: int, world: int, dev: str) -> dict:
    fx = load_golden(name)
    r = fx["recipe"]
    tr = _trainer(r, False, dev, data_parallel=True)
    x_all = O.synthetic_activations(r["B"] * r["steps"], r["d"], r["data_seed"])
    worst_loss, strict = 0.0, True
    for s in range(r["steps"]):
        a, b = parallel.shard_rows(r["B"], world, rank)
        m = tr.train_step(x_all[s * r["B"] + a:s * r["B"] + b].to(dev))
        ref = fx["per_step"][s]
        if ref["min_gap_rel"] <= 2e-6:
            strict = False
        worst_loss = max(worst_loss, abs(m.loss - ref["loss"]) / abs(ref["loss"]))
        assert m.l0 == ref["l0"] or not strict, (m.l0, ref["l0"])
        assert abs(m.dead_feature_ratio - ref["dead_feature_ratio"]) < 1e-7 or not strict
        assert abs(m.learning_rate - ref["lr_reported"]) <= 1e-9 * ref["lr_reported"]
    tr.consolidate_weights()
    sae = tr.model
    counters_equal = bool(torch.equal(sae.feature_last_activated.cpu(),
                                      fx["final_counters"]["feature_last_activated"]))
    worst_w = worst_l2 = 0.0
    sd = sae.state_dict()
    for n in O.PARAM_ORDER:
        ref = fx["final_params"][n]
        t = sd[n].cpu()
        if isinstance(ref, dict):
            got, want = t.contiguous().reshape(-1)[:: ref["sample_stride"]], ref["sample"]
        else:
            got, want = t, ref
        worst_w = max(worst_w, ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item())
        worst_l2 = max(worst_l2, ((got.double() - want.double()).norm() / want.double().norm().clamp_min(1e-300)).item())
    return {"case": name, "loss_rel": worst_loss, "weight_rel": worst_w, "weight_rel_l2": worst_l2,
            "counters_equal": counters_equal,
            "strict": strict, "step_count": int(sae.step_count), "mode": str(tr.cuda_graph)}


def single_vs_sharded_bf16(rank: int, world: int, dev: str) -> dict:
    r = dict(d=768, F=6144, k=32, B=4096, steps=4, total_steps=100, lr=1e-3, warmup=2, dead_threshold=1,
             model_seed=3)
    x = O.synthetic_activations(r["B"] * r["steps"], r["d"], seed=11).to(dev)
    single = _trainer(r, True, dev)
    ref = [single.train_step(x[s * r["B"]:(s + 1) * r["B"]]) for s in range(r["steps"])]
    dp = _trainer(r, True, dev, data_parallel=True)
    got = []
    for s in range(r["steps"]):
        a, b = parallel.shard_rows(r["B"], world, rank)
        got.append(dp.train_step(x[s * r["B"] + a:s * r["B"] + b]))
    dp.consolidate_weights()      # bf16 operand gather: fp32 rows of the other ranks are gathered on demand
    loss_rel = max(abs(g.loss - q.loss) / abs(q.loss) for g, q in zip(got, ref))
    l0_equal = all(g.l0 == q.l0 for g, q in zip(got, ref))
    dead_equal = all(g.dead_feature_ratio == q.dead_feature_ratio for g, q in zip(got, ref))
    worst = 0.0
    for (n, p), (_, q) in zip(dp.model.named_parameters(), single.model.named_parameters()):
        worst = max(worst, ((p - q).abs().max() / q.abs().max().clamp_min(1e-30)).item())
    # replicas must stay bit-identical without a weight broadcast: compare a checksum over the ranks
    chk = torch.stack([p.detach().double().sum() for p in dp.model.parameters()])
    lo, hi = chk.clone(), chk.clone()
    torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
    torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
    return {"case": "bf16_768x6144_vs_single", "loss_rel": loss_rel, "l0_equal": l0_equal,
            "dead_equal": dead_equal, "weight_rel": worst,
            "counters_equal": bool(torch.equal(dp.model.feature_last_activated,
                                               single.model.feature_last_activated)),
            "replicas_identical": bool(torch.equal(lo, hi)), "mode": str(dp.cuda_graph)}


def main() -> None:
    out = Path(sys.argv[1])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = []
    try:
        for name in ("tiny_test_384x3072", "small_768x6144", "large_1280x40960"):
            res.append(golden_fp32(name, rank, world, dev))
        res.append(single_vs_sharded_bf16(rank, world, dev))
    finally:
        (out / f"rank{rank}.json").write_text(json.dumps(res, indent=1))
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
