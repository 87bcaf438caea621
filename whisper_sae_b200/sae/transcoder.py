"""Transcoders with the reference's surface (``whisper_sae.sae.transcoder``,
/root/reference/src/whisper_sae/sae/transcoder.py:21-461) on the fused sm_100a TopK path.

A transcoder is the TopK-SAE hot path with a target that differs from the input and no ``b_pre``:
``pred = TopK(W_enc x + b_enc) W_dec^T + b_dec`` (:105-138), loss = MSE(pred, mlp_output) (:149).
The same autograd node (``_FusedTopKSAE``: K0 pack, K1 tcgen05 GEMM + TopK, K2 sparse decode,
K3/K4 backward) serves both.  ``SkipTranscoder`` (:244-422) adds a dense affine skip path; since
``mse(sparse + skip, y) == mse(sparse, y - skip)`` the skip output is folded into the target of the
fused node and its gradient flows back through ``d loss / d target`` (the skip GEMM is genuinely
dense: library ``nn.Linear``).
"""

from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import ops
from .model import _feature_major_, _FusedTopKSAE, _LazyOutput, _SparseState, _fp32_terms, _warn_encode_detached


class TranscoderOutput(_LazyOutput):
    """``predicted, hidden, loss, reconstruction_loss, sparsity_loss, l0`` (transcoder.py:21-29)."""

    _fields = ("predicted", "hidden", "loss", "reconstruction_loss", "sparsity_loss", "l0")
    __slots__ = ()


class _TranscoderBase(nn.Module):
    def __init__(self, input_dim: int, output_dim: int, hidden_dim: int, k: int,
                 normalize_decoder: bool, dead_feature_threshold: int, precision: str | None):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.hidden_dim = hidden_dim
        self.k = k
        self.normalize_decoder = normalize_decoder
        self.dead_feature_threshold = dead_feature_threshold
        if precision not in (None, "bf16", "fp32"):
            raise ValueError("precision must be None, 'bf16' or 'fp32'")
        self.precision = precision
        self.encoder = nn.Linear(input_dim, hidden_dim, bias=True)
        self.decoder = nn.Linear(hidden_dim, output_dim, bias=True)
        self._global_rows: int | None = None

    def _finish_init(self) -> None:
        self.register_buffer("feature_last_activated", torch.zeros(self.hidden_dim, dtype=torch.long))
        self.register_buffer("step_count", torch.tensor(0, dtype=torch.long))
        _feature_major_(self.decoder)

    # ---- shared helpers -------------------------------------------------------------------
    def _use_bf16(self) -> bool:
        if self.precision is not None:
            return self.precision == "bf16"
        return torch.is_autocast_enabled("cuda")

    def _check(self, x: Tensor, dim: int, what: str) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError("whisper_sae_b200 transcoders run on CUDA sm_100a only (no CPU fallback)")
        if x.dim() != 2 or x.shape[1] != dim:
            raise RuntimeError(f"expected {what} of shape [batch, {dim}], got {tuple(x.shape)}")
        return x

    def _w_decT(self) -> Tensor:
        _feature_major_(self.decoder)
        return self.decoder.weight.data.t()

    def normalize_decoder_weights(self) -> None:
        """Unit-norm decoder columns, in place (transcoder.py:95-100 / :321-326); zero columns stay
        zero (1e-12 clamp), which is what SkipTranscoder's zero-initialised decoder relies on."""
        w = self._w_decT()
        if w.is_cuda:
            ops.renorm_decoder_(w, 1e-12)
        else:
            with torch.no_grad():
                w.copy_(nn.functional.normalize(w, dim=1))

    def encode(self, x: Tensor) -> Tensor:
        """Dense [batch, hidden_dim] TopK activations (transcoder.py:102-118). Not differentiable."""
        x = self._check(x, self.input_dim, "mlp_input")
        if torch.is_grad_enabled() and (x.requires_grad or self.encoder.weight.requires_grad):
            _warn_encode_detached()       # no grad_fn on the result (see TopKSAE.encode)
        terms = 1 if self._use_bf16() else _fp32_terms()
        x32 = x.detach().to(torch.float32).contiguous()
        a = ops.pack_activations(x32, None, terms)
        w = ops.pack_encoder(self.encoder.weight.detach().contiguous(), self.encoder.bias.detach(), terms)
        idx, val = ops.encode_topk(a, w, x.shape[0], self.hidden_dim, self.input_dim, terms, self.k)
        return ops.densify_hidden(idx, val, self.hidden_dim)

    def decode(self, hidden: Tensor) -> Tensor:
        return nn.functional.linear(hidden, self.decoder.weight, self.decoder.bias)

    def _fused(self, mlp_input: Tensor, target: Tensor) -> tuple[Tensor, _SparseState]:
        self._w_decT()
        st = _SparseState()
        loss = _FusedTopKSAE.apply(self, st, self._use_bf16(), mlp_input, target, None,
                                   self.encoder.weight, self.encoder.bias, self.decoder.weight,
                                   self.decoder.bias)
        self._last_sparse = st
        return loss, st

    def _update_dead_features(self, hidden: Tensor) -> None:
        if self.training:
            self.step_count += 1
            self.feature_last_activated[(hidden > 0).any(dim=0)] = self.step_count

    def get_dead_features(self) -> Tensor:
        return (self.step_count - self.feature_last_activated) > self.dead_feature_threshold

    def get_dead_feature_ratio(self) -> float:
        return self.get_dead_features().float().mean().item()


class TopKTranscoder(_TranscoderBase):
    """TopK transcoder (transcoder.py:32-241), fused sm_100a implementation."""

    def __init__(self, input_dim: int, output_dim: int, hidden_dim: int, k: int = 32,
                 normalize_decoder: bool = True, dead_feature_threshold: int = 10_000,
                 precision: str | None = None):
        super().__init__(input_dim, output_dim, hidden_dim, k, normalize_decoder,
                         dead_feature_threshold, precision)
        with torch.no_grad():          # xavier-uniform, unit-norm columns, x0.1 (transcoder.py:86-93)
            nn.init.xavier_uniform_(self.decoder.weight)
            self.decoder.weight.data = nn.functional.normalize(self.decoder.weight.data, dim=0)
            self.decoder.weight.data *= 0.1
        self._finish_init()

    def forward(self, mlp_input: Tensor, mlp_output: Tensor) -> TranscoderOutput:
        x = self._check(mlp_input, self.input_dim, "mlp_input")
        y = self._check(mlp_output, self.output_dim, "mlp_output")
        loss, st = self._fused(x, y)
        B, F = x.shape[0], self.hidden_dim
        y32 = y.detach().to(torch.float32)
        return TranscoderOutput(
            lambda: st.resid + y32,
            lambda: ops.densify_hidden(st.idx, st.val, F),
            loss, loss,
            lambda: torch.zeros((), dtype=torch.float32, device=x.device),
            lambda: st.stats[1].to(torch.float32) / float(B))

    def resample_dead_features(self, mlp_inputs: Tensor, mlp_outputs: Tensor,
                               num_resample: int | None = None) -> int:
        """transcoder.py:188-241, vectorised on device: the i-th dead feature (ascending) gets the
        i-th highest-error input as encoder row and the normalised residual as decoder column."""
        dead_idx = torch.where(self.get_dead_features())[0]
        num_dead = int(dead_idx.numel())
        if num_dead == 0:
            return 0
        if num_resample is not None:
            num_dead = min(num_dead, num_resample)
            dead_idx = dead_idx[:num_dead]
        with torch.no_grad():
            self.forward(mlp_inputs, mlp_outputs)
            residuals = -self._last_sparse.resid            # mlp_outputs - predicted
            errors = (residuals ** 2).sum(dim=-1)
            n = min(num_dead, errors.numel())
            _, top = torch.topk(errors, n)
            tgt = dead_idx[:n]
            self.encoder.weight.data[tgt] = nn.functional.normalize(mlp_inputs[top].float(), dim=-1)
            self.encoder.bias.data[tgt] = 0.0
            self._w_decT()[tgt] = nn.functional.normalize(residuals[top], dim=-1)
            self.feature_last_activated[tgt] = self.step_count
        return num_dead


class SkipTranscoder(_TranscoderBase):
    """TopK transcoder + affine skip path (transcoder.py:244-422)."""

    def __init__(self, input_dim: int, output_dim: int, hidden_dim: int, k: int = 32,
                 normalize_decoder: bool = True, dead_feature_threshold: int = 10_000,
                 precision: str | None = None):
        super().__init__(input_dim, output_dim, hidden_dim, k, normalize_decoder,
                         dead_feature_threshold, precision)
        self.skip = nn.Linear(input_dim, output_dim, bias=True)
        with torch.no_grad():          # decoder and skip start at zero (transcoder.py:300-319)
            nn.init.zeros_(self.decoder.weight)
            nn.init.zeros_(self.decoder.bias)
            nn.init.zeros_(self.skip.weight)
            nn.init.zeros_(self.skip.bias)
        self._finish_init()

    def set_output_bias(self, mean_output: Tensor) -> None:
        with torch.no_grad():
            self.decoder.bias.data = mean_output.clone()

    def forward(self, mlp_input: Tensor, mlp_output: Tensor) -> TranscoderOutput:
        x = self._check(mlp_input, self.input_dim, "mlp_input")
        y = self._check(mlp_output, self.output_dim, "mlp_output")
        skip_out = self.skip(x.to(self.skip.weight.dtype)).to(torch.float32)
        # mse(sparse + skip, y) == mse(sparse, y - skip): the skip path enters through the target
        loss, st = self._fused(x, y.to(torch.float32) - skip_out)
        B, F = x.shape[0], self.hidden_dim
        y32 = y.detach().to(torch.float32)
        return TranscoderOutput(
            lambda: st.resid + y32,
            lambda: ops.densify_hidden(st.idx, st.val, F),
            loss, loss,
            lambda: torch.zeros((), dtype=torch.float32, device=x.device),
            lambda: st.stats[1].to(torch.float32) / float(B))

    def get_skip_contribution(self, mlp_input: Tensor, mlp_output: Tensor) -> float:
        with torch.no_grad():
            skip_var = ((self.skip(mlp_input) - mlp_output) ** 2).mean()
            total_var = ((mlp_output - mlp_output.mean(dim=0)) ** 2).mean()
            return (1 - (skip_var / (total_var + 1e-8))).item()


def create_transcoder(input_dim: int, output_dim: int, hidden_dim: int, k: int = 32,
                      use_skip: bool = True, **kwargs) -> nn.Module:
    """Factory (transcoder.py:425-461)."""
    cls = SkipTranscoder if use_skip else TopKTranscoder
    return cls(input_dim=input_dim, output_dim=output_dim, hidden_dim=hidden_dim, k=k, **kwargs)
