"""SAE trainer with the reference's interface (``whisper_sae.sae.training``,
/root/reference/src/whisper_sae/sae/training.py:19-379) over the fused sm_100a train step.

Same constructor, public attributes, ``train_step`` / ``train_epoch`` / ``train`` /
``setup_scheduler`` / checkpoint + metrics file formats.  Differences that are deliberate:

* the five per-step ``.item()`` host syncs (training.py:207-215) become ONE 24-byte readback of
  the packed stats buffer written by the kernels (SSE as float64, L0 count, dead count);
* ``use_amp`` selects the bf16 tensor-core path instead of fp16 autocast.  bf16 keeps the fp32
  exponent range, so loss scaling is the identity: ``self.scaler`` is still a ``GradScaler``
  (attribute preserved) but constructed disabled unless ``grad_scaler=True`` is passed; the fused
  autograd node honours an arbitrary scalar ``grad_output`` either way;
* ``fused_optimizer=True`` (default on CUDA) runs clip-by-global-norm + AdamW as one pass per
  parameter over the optimizer's own state tensors (csrc/wsae_elementwise.cu), so
  ``optimizer.state_dict()`` checkpoints stay interchangeable with ``torch.optim.AdamW``.
"""

from __future__ import annotations

import json
import math
from dataclasses import asdict, dataclass
from pathlib import Path

import torch
from rich.progress import BarColumn, Progress, SpinnerColumn, TaskProgressColumn, TextColumn
from torch import Tensor
from torch.optim import AdamW
from torch.optim.lr_scheduler import CosineAnnealingLR, LinearLR, SequentialLR
from torch.utils.data import DataLoader

from .. import ops
from ..config import TrainingConfig
from .model import SAEOutput, TopKSAE


@dataclass
class TrainingMetrics:
    """Per-step metrics (training.py:19-29)."""

    loss: float
    reconstruction_loss: float
    sparsity_loss: float
    l0: float
    dead_feature_ratio: float
    learning_rate: float
    step: int


class SAETrainer:
    """Trainer for sparse autoencoders (training.py:32-379)."""

    def __init__(
        self,
        model: TopKSAE,
        config: TrainingConfig,
        device: torch.device | str = "cpu",
        run_dir: Path | None = None,
        resample_dead_every: int = 5000,
        resample_batch_size: int = 8192,
        *,
        grad_scaler: bool = False,
        fused_optimizer: bool | None = None,
    ):
        self.model = model.to(device)
        self.config = config
        self.device = device
        self.run_dir = run_dir or Path("outputs")
        self.run_dir.mkdir(parents=True, exist_ok=True)
        self.resample_dead_every = resample_dead_every
        self.resample_batch_size = resample_batch_size

        self.optimizer = AdamW(model.parameters(), lr=config.learning_rate,
                               weight_decay=config.weight_decay)
        self.scheduler = None

        is_cuda = str(device).startswith("cuda")
        self.use_amp = config.use_amp and is_cuda
        self.scaler = torch.amp.GradScaler("cuda", enabled=self.use_amp and grad_scaler)
        self.fused_optimizer = is_cuda if fused_optimizer is None else (fused_optimizer and is_cuda)

        self.global_step = 0
        self.epoch = 0
        self.metrics_history: list[TrainingMetrics] = []
        self.num_resampled_total = 0
        self.wandb_run = None
        self._resample_dataset = None
        self._hyper: Tensor | None = None
        self._sumsq: Tensor | None = None

    # ------------------------------------------------------------------ resampling plumbing
    def set_resample_dataset(self, dataset: torch.utils.data.Dataset) -> None:
        self._resample_dataset = dataset

    def _maybe_resample_dead_features(self) -> int:
        """Same gating as training.py:97-134 (the reference never calls it automatically either)."""
        if self._resample_dataset is None or not hasattr(self.model, "resample_dead_features"):
            return 0
        if self.global_step == 0 or self.global_step % self.resample_dead_every != 0:
            return 0
        picks = torch.randperm(len(self._resample_dataset))[: self.resample_batch_size]
        tensors = getattr(self._resample_dataset, "tensors", None)
        if tensors is not None:  # TensorDataset: one indexed gather instead of a Python loop
            batch = tensors[0][picks]
        else:
            rows = [self._resample_dataset[int(i)] for i in picks]
            rows = [r[0] if isinstance(r, tuple) else r for r in rows]
            batch = torch.stack(rows)
        n = self.model.resample_dead_features(batch.to(self.device))
        self.num_resampled_total += n
        if n > 0 and self.wandb_run is not None:
            self.wandb_run.log({"train/features_resampled": n}, step=self.global_step)
        return n

    # ------------------------------------------------------------------ schedule
    def setup_scheduler(self, total_steps: int) -> None:
        """Linear warm-up (0.01 -> 1) then cosine to 0.1 * lr (training.py:136-159)."""
        warmup = min(self.config.warmup_steps, total_steps // 10)
        ramp = LinearLR(self.optimizer, start_factor=0.01, end_factor=1.0, total_iters=warmup)
        decay = CosineAnnealingLR(self.optimizer, T_max=total_steps - warmup,
                                  eta_min=self.config.learning_rate * 0.1)
        self.scheduler = SequentialLR(self.optimizer, schedulers=[ramp, decay], milestones=[warmup])

    # ------------------------------------------------------------------ optimizer step
    def _fused_clip_adamw(self) -> None:
        """clip_grad_norm_ + AdamW.step in len(params)+len(params) kernel launches, no host sync."""
        group = self.optimizer.param_groups[0]
        params = [p for p in group["params"] if p.grad is not None]
        if not params:
            return
        dev = params[0].device
        if self._sumsq is None:
            self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
            self._hyper = torch.empty(8, dtype=torch.float32, device=dev)
        self._sumsq.zero_()
        for p in params:
            g = p.grad
            if not (g.is_contiguous() or g.t().is_contiguous()):
                p.grad = g = g.contiguous()
            ops.sumsq_(g, self._sumsq)
        # optimizer state, created exactly like torch.optim.AdamW does (step as a float32 tensor)
        step_t = None
        for p in params:
            state = self.optimizer.state[p]
            if len(state) == 0:
                state["step"] = torch.tensor(0.0, dtype=torch.float32)
                state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            state["step"] += 1
            step_t = float(state["step"])
        beta1, beta2 = group["betas"]
        lr = group["lr"]
        bc1 = 1.0 - beta1 ** step_t
        bc2s = math.sqrt(1.0 - beta2 ** step_t)
        hyper = torch.tensor([lr, beta1, beta2, group["eps"], group["weight_decay"], bc1, bc2s,
                              self.config.gradient_clip], dtype=torch.float32)
        self._hyper.copy_(hyper, non_blocking=True)
        for p in params:
            state = self.optimizer.state[p]
            # parameter, grad and moments share one dense layout (contiguous or its transpose),
            # so a flat elementwise pass over the raw storage is exact
            if p.grad.stride() != p.stride():
                same_layout = torch.empty_strided(p.shape, p.stride(), dtype=p.dtype, device=p.device)
                p.grad = same_layout.copy_(p.grad)
            ops.fused_adamw_(p.data, p.grad, state["exp_avg"], state["exp_avg_sq"], self._hyper,
                             self._sumsq)

    # ------------------------------------------------------------------ the hot loop
    def train_step(self, batch: Tensor | tuple | list) -> TrainingMetrics:
        """One optimisation step; same order of operations as training.py:161-217."""
        self.model.train()
        if isinstance(batch, (tuple, list)):
            batch = batch[0]
        batch = batch.to(self.device, non_blocking=True)

        with torch.amp.autocast("cuda", enabled=self.use_amp):
            output: SAEOutput = self.model(batch)

        self.optimizer.zero_grad()
        self.scaler.scale(output.loss).backward()

        if self.fused_optimizer and not self.scaler.is_enabled():
            self._fused_clip_adamw()
        else:
            self.scaler.unscale_(self.optimizer)
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.config.gradient_clip)
            self.scaler.step(self.optimizer)
            self.scaler.update()

        if hasattr(self.model, "normalize_decoder_weights"):
            self.model.normalize_decoder_weights()
        if self.scheduler is not None:
            self.scheduler.step()
        self.global_step += 1

        metrics = self._read_metrics(output, batch.shape[0])
        return metrics

    def _read_metrics(self, output: SAEOutput, rows: int) -> TrainingMetrics:
        lr = self.optimizer.param_groups[0]["lr"]
        st = getattr(self.model, "_last_sparse", None)
        if st is not None and st.stats is not None and output.loss.is_cuda:
            raw = st.stats.cpu()  # the step's single device->host sync (24 bytes)
            sse = raw[:1].view(torch.float64).item()
            d_out = st.resid.shape[1]
            loss = float(torch.tensor(sse / (float(st.rows_total) * d_out), dtype=torch.float32))
            l0 = float(torch.tensor(raw[1].item() / float(rows), dtype=torch.float32))
            hidden_dim = self.model.feature_last_activated.numel()
            dead = float(torch.tensor(raw[2].item(), dtype=torch.float32) / hidden_dim)
            return TrainingMetrics(loss, loss, 0.0, l0, dead, lr, self.global_step)
        return TrainingMetrics(
            loss=output.loss.item(),
            reconstruction_loss=output.reconstruction_loss.item(),
            sparsity_loss=output.sparsity_loss.item(),
            l0=output.l0.item(),
            dead_feature_ratio=self.model.get_dead_feature_ratio(),
            learning_rate=lr,
            step=self.global_step,
        )

    def train_epoch(self, dataloader: DataLoader, progress: Progress | None = None,
                    task_id: int | None = None) -> list[TrainingMetrics]:
        epoch_metrics: list[TrainingMetrics] = []
        for batch in dataloader:
            m = self.train_step(batch)
            epoch_metrics.append(m)
            self.metrics_history.append(m)
            if progress is not None and task_id is not None:
                progress.update(task_id, advance=1)
            if self.wandb_run is not None and self.global_step % 100 == 0:
                self.wandb_run.log(
                    {"train/loss": m.loss, "train/reconstruction_loss": m.reconstruction_loss,
                     "train/l0": m.l0, "train/dead_ratio": m.dead_feature_ratio,
                     "train/lr": m.learning_rate},
                    step=self.global_step,
                )
        self.epoch += 1
        return epoch_metrics

    def train(self, dataloader: DataLoader, epochs: int | None = None,
              checkpoint_every: int | None = None) -> None:
        epochs = epochs or self.config.epochs
        checkpoint_every = checkpoint_every or self.config.checkpoint_every
        self.setup_scheduler(len(dataloader) * epochs)
        columns = (SpinnerColumn(), TextColumn("[progress.description]{task.description}"),
                   BarColumn(), TaskProgressColumn())
        with Progress(*columns) as progress:
            outer = progress.add_task(f"[cyan]Training {epochs} epochs", total=epochs)
            for e in range(epochs):
                inner = progress.add_task(f"[green]Epoch {e + 1}/{epochs}", total=len(dataloader))
                ms = self.train_epoch(dataloader, progress, inner)
                mean_loss = sum(m.loss for m in ms) / len(ms)
                mean_l0 = sum(m.l0 for m in ms) / len(ms)
                progress.remove_task(inner)
                progress.update(outer, advance=1)
                progress.console.print(
                    f"Epoch {e + 1}: loss={mean_loss:.4f}, L0={mean_l0:.1f}, "
                    f"dead={ms[-1].dead_feature_ratio:.1%}"
                )
                if (e + 1) % checkpoint_every == 0:
                    self.save_checkpoint(f"checkpoint_epoch{e + 1}.pt")
        self.save_checkpoint("final.pt")

    # ------------------------------------------------------------------ persistence
    def save_checkpoint(self, filename: str) -> Path:
        """Same dict layout as training.py:328-338."""
        path = self.run_dir / filename
        payload = {
            "model_state_dict": self.model.state_dict(),
            "optimizer_state_dict": self.optimizer.state_dict(),
            "scheduler_state_dict": self.scheduler.state_dict() if self.scheduler else None,
            "global_step": self.global_step,
            "epoch": self.epoch,
            "config": self.config.model_dump(),
        }
        torch.save(payload, path)
        return path

    def load_checkpoint(self, path: str | Path) -> None:
        ckpt = torch.load(path, map_location=self.device)
        self.model.load_state_dict(ckpt["model_state_dict"])
        self.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        if ckpt["scheduler_state_dict"] and self.scheduler:
            self.scheduler.load_state_dict(ckpt["scheduler_state_dict"])
        self.global_step = ckpt["global_step"]
        self.epoch = ckpt["epoch"]

    def save_metrics(self, filename: str = "metrics.json") -> Path:
        path = self.run_dir / filename
        keys = ("step", "loss", "reconstruction_loss", "sparsity_loss", "l0",
                "dead_feature_ratio", "learning_rate")
        rows = [{k: asdict(m)[k] for k in keys} for m in self.metrics_history]
        path.write_text(json.dumps(rows, indent=2))
        return path
