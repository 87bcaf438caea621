#!/usr/bin/env python
"""Train one TopK-SAE per (component, layer) on cached activations — the recipe of the reference's
``scripts/train.py::train_layer`` (/root/reference/scripts/train.py:118-219) on the B200 path, with
the layers dealt over the GPUs of one box (BASELINE config 2: whisper-tiny, 4 encoder + 4 decoder
layers, one layer per B200, no collective).

    python scripts/train_layers.py --config configs/tiny_default.yaml --synthetic-rows 1048576
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/train_layers.py ...

Whisper checkpoints and LibriSpeech are not available offline, so ``--synthetic-rows N`` writes
row-standardised Gaussian activations (what the reference's extraction yields from a random-init
Whisper: the hooks apply the final LayerNorm) through ``FeatureCache.save`` for every
layer that has no cache yet; everything after that is the real load -> dataloader -> trainer ->
``sae_final.pt`` + ``metrics.json`` path.
"""

from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from whisper_sae_b200.config import ExperimentConfig  # noqa: E402
from whisper_sae_b200.data import FeatureCache  # noqa: E402
from whisper_sae_b200.sae import SAETrainer, create_sae  # noqa: E402
from whisper_sae_b200.sae.parallel import layer_assignment  # noqa: E402


def synthetic_activations(n_rows: int, d: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_rows, d, generator=g)
    return (x - x.mean(1, keepdim=True)) / x.std(1, unbiased=False, keepdim=True)


def train_layer(config: ExperimentConfig, component: str, layer_idx: int, cache: FeatureCache,
                device: str, resident: bool, epochs: int | None) -> dict:
    """scripts/train.py:118-219: load cache -> create_sae -> dataloader -> SAETrainer.train ->
    sae_final.pt (bare state_dict) + metrics.json in outputs/{experiment}_{component}_layer{idx}."""
    features, meta = cache.load(component, layer_idx)
    input_dim = features.shape[-1]
    torch.manual_seed(config.training.seed)
    sae = create_sae(config.sae, input_dim)
    loader = cache.get_dataloader(component, layer_idx, config.training.batch_size, shuffle=True,
                                  num_workers=config.training.num_workers,
                                  device=device if resident else None)
    run_dir = config.output_dir / f"{config.experiment_name}_{component}_layer{layer_idx}"
    trainer = SAETrainer(sae, config.training, device=device, run_dir=run_dir)
    trainer.set_resample_dataset(torch.utils.data.TensorDataset(features))
    trainer.train(loader, epochs=epochs)
    torch.save(sae.state_dict(), run_dir / "sae_final.pt")
    trainer.save_metrics()
    last = trainer.metrics_history[-1]
    return {"component": component, "layer": layer_idx, "steps": trainer.global_step, "loss": last.loss,
            "l0": last.l0, "dead": last.dead_feature_ratio, "run_dir": str(run_dir), "rows": meta.num_tokens}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=Path, default=ROOT / "configs" / "tiny_default.yaml")
    ap.add_argument("--synthetic-rows", type=int, default=0)
    ap.add_argument("--epochs", type=int, default=None)
    ap.add_argument("--batch-size", type=int, default=None)
    ap.add_argument("--output-dir", type=Path, default=None)
    ap.add_argument("--cache-dir", type=Path, default=None)
    ap.add_argument("--host-loader", action="store_true",
                    help="use the reference's DataLoader(TensorDataset) instead of resident batches")
    args = ap.parse_args()

    cfg = ExperimentConfig.from_yaml(args.config)
    if args.batch_size:
        cfg.training.batch_size = args.batch_size
    if args.output_dir:
        cfg.output_dir = args.output_dir
    if args.cache_dir:
        cfg.data.cache_dir = args.cache_dir
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("train_layers.py needs a CUDA device (B200): the SAE path has no CPU fallback")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    cache = FeatureCache(Path(cfg.data.cache_dir) / "features", cfg.whisper, cfg.data)   # scripts/train.py:275
    units = layer_assignment(list(cfg.encoder_layers), list(cfg.decoder_layers), world, rank)
    for component, layer in units:
        if not cache.has_cache(component, layer):
            if args.synthetic_rows <= 0:
                raise SystemExit(f"no cache for {component} layer {layer} (pass --synthetic-rows N)")
            seed = 1234 + layer + (100 if component == "decoder" else 0)
            cache.save(synthetic_activations(args.synthetic_rows, cfg.whisper.hidden_dim, seed),
                       component, layer, num_samples=args.synthetic_rows)
    for component, layer in units:
        out = train_layer(cfg, component, layer, cache, device, not args.host_loader, args.epochs)
        print(f"[rank {rank}] {component} layer {layer}: {out['steps']} steps over {out['rows']} rows, "
              f"loss {out['loss']:.5f}, L0 {out['l0']:.1f}, dead {out['dead']:.1%} -> {out['run_dir']}")


if __name__ == "__main__":
    main()
