"""Whole-step CUDA graph for the variant modules (TopK crosscoder, TopK / skip transcoder).

The reference has no trainer for these: its tests and notebooks step them by hand
(`tests/test_crosscoder.py:429-470`, `tests/test_transcoder.py:140-175` in the reference tree):

    out = model(inputs); optimizer.zero_grad(); out.loss.backward(); optimizer.step()
    model.normalize_decoder_weights()

Launched like that a step at 16 384 rows is ~25 kernels of 5-150 us each behind ~0.8 ms of Python and
autograd dispatch.  :class:`GraphedVariantStep` runs exactly that sequence - the same modules, the
same autograd node over K0 / K1 / K23 / K4, the caller's own torch optimizer - but captured once into
one CUDA graph and replayed: the host cost of a step becomes one copy of the inputs into the graph's
static buffers plus one graph launch.  ``SAETrainer`` keeps its own hand-built graph for ``TopKSAE``;
this is the generic form for every module whose forward returns an object with ``.loss``.

The first call runs eagerly (it creates the optimizer state and the lazily built layouts), the
second call captures and replays, later calls replay.  Inputs must keep their shapes; a new shape
re-captures.  The optimizer must be capturable (``torch.optim.AdamW(..., capturable=True)`` - its
step counters live on the device); `make_optimizer` builds one.
"""

from __future__ import annotations

import torch
from torch import Tensor, nn

from .training import _quiet_gc

__all__ = ["GraphedVariantStep", "make_optimizer"]


def make_optimizer(model: nn.Module, lr: float = 1e-4, weight_decay: float = 0.0) -> torch.optim.AdamW:
    """AdamW with the reference trainer's hyper-parameters (training.py:63-67) in its capturable form."""
    return torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=weight_decay, capturable=True,
                             fused=True)


def _static_like(v):
    if isinstance(v, Tensor):
        return torch.empty_like(v, memory_format=torch.contiguous_format)
    if isinstance(v, dict):
        return {k: _static_like(t) for k, t in v.items()}
    raise TypeError(f"GraphedVariantStep inputs must be tensors or dicts of tensors, got {type(v).__name__}")


def _signature(v):
    if isinstance(v, Tensor):
        return (tuple(v.shape), v.dtype)
    return tuple((k, tuple(t.shape), t.dtype) for k, t in v.items())


def _fill(dst, src) -> None:
    if isinstance(dst, Tensor):
        dst.copy_(src, non_blocking=True)
    else:
        for k, t in dst.items():
            t.copy_(src[k], non_blocking=True)


class StepResult:
    """What a graphed step hands back: device tensors, no synchronisation.  ``loss`` / ``l0`` are
    snapshots (clones) of the graph's static outputs, so they stay valid after the next replay."""

    __slots__ = ("loss", "l0")

    def __init__(self, loss: Tensor, l0: Tensor):
        self.loss, self.l0 = loss, l0


class GraphedVariantStep:
    def __init__(self, model: nn.Module, optimizer: torch.optim.Optimizer, *, use_amp: bool = True,
                 gradient_clip: float | None = None):
        if not next(model.parameters()).is_cuda:
            raise RuntimeError("GraphedVariantStep runs on CUDA sm_100a only (no CPU fallback)")
        for group in optimizer.param_groups:
            if not group.get("capturable", False):
                raise RuntimeError("the optimizer must be built with capturable=True (see make_optimizer)")
        self.model, self.optimizer = model, optimizer
        self.use_amp = bool(use_amp)
        self.gradient_clip = gradient_clip
        self.calls = 0
        self._sig = None
        self._graph: torch.cuda.CUDAGraph | None = None
        self._static: tuple = ()
        self._out = None

    def _body(self, inputs: tuple):
        # grads set to None first: autograd then ASSIGNS the freshly computed gradient tensors instead of
        # adding them into the old ones (one read-modify-write pass over every parameter less; inside the
        # captured graph the assigned tensors live at fixed pool addresses, so replays stay valid)
        self.optimizer.zero_grad(set_to_none=True)
        with torch.amp.autocast("cuda", enabled=self.use_amp, dtype=torch.bfloat16):
            out = self.model(*inputs)
        out.loss.backward()
        if self.gradient_clip is not None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.gradient_clip, foreach=True)
        self.optimizer.step()
        if hasattr(self.model, "normalize_decoder_weights"):
            self.model.normalize_decoder_weights()
        return out

    def _result(self, out, rows: int) -> StepResult:
        st = getattr(self.model, "_last_sparse", None)
        if st is not None and st.stats is not None:
            l0 = st.stats[1].to(torch.float32) / float(rows)
        else:
            l0 = out.l0.detach().clone()
        return StepResult(out.loss.detach().clone(), l0)

    def __call__(self, *inputs) -> StepResult:
        self.model.train()
        first = inputs[0] if isinstance(inputs[0], Tensor) else next(iter(inputs[0].values()))
        rows = first.shape[0]
        sig = tuple(_signature(v) for v in inputs)
        if sig != self._sig:          # new shapes: start over (eager step, then capture)
            self._sig, self._graph, self.calls = sig, None, 0
        self.calls += 1
        if self.calls == 1:
            return self._result(self._body(inputs), rows)
        if self._graph is None:
            self._static = tuple(_static_like(v) for v in inputs)
            for dst, src in zip(self._static, inputs):
                _fill(dst, src)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with _quiet_gc(), torch.cuda.graph(g):
                self._out = self._body(self._static)
            self._graph = g
        else:
            for dst, src in zip(self._static, inputs):
                _fill(dst, src)
        self._graph.replay()
        return self._result(self._out, rows)
