"""Cross-layer crosscoders with the reference's surface (``whisper_sae.sae.crosscoder``,
/root/reference/src/whisper_sae/sae/crosscoder.py:26-417).

``TopKCrossLayerCrosscoder`` (:298-379) is the TopK-SAE hot path over the concatenation of the
per-layer activations: ``pre = sum_l x_l W_enc[l] + b_enc`` is one GEMM with K = L*d
(``x_cat [B, L*d]`` against ``W_enc`` viewed as ``[F, L*d]``), the shared decoder ``W_dec [F, L, d]``
is already feature-major (``[F, L*d]`` rows), and the loss ``sum_l mean((recon_l - x_l)^2)`` is
``L * mean`` over the concatenated residual.  It runs on the same fused autograd node as
``TopKSAE``.  The ReLU + L1 ``CrossLayerCrosscoder`` (:38-296) is genuinely dense and keeps
PyTorch math.
"""

from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import ops
from .model import _FusedTopKSAE, _LazyOutput, _SparseState, _fp32_terms, _warn_encode_detached


class CrosscoderOutput(_LazyOutput):
    """``reconstructed{layer}, hidden, loss, reconstruction_loss, sparsity_loss, l0,
    per_layer_loss{layer}`` (crosscoder.py:26-35)."""

    _fields = ("reconstructed", "hidden", "loss", "reconstruction_loss", "sparsity_loss", "l0",
               "per_layer_loss")
    __slots__ = ()


class CrossLayerCrosscoder(nn.Module):
    """ReLU + decoder-norm-weighted L1 crosscoder (crosscoder.py:38-296); dense => library math."""

    def __init__(self, d_model: int, n_layers: int, d_sae: int, layer_indices: list[int] | None = None,
                 activation: str = "relu", sparsity_weight: float = 0.01,
                 normalize_decoder: bool = True, dead_feature_threshold: int = 10_000):
        super().__init__()
        self.d_model = d_model
        self.n_layers = n_layers
        self.d_sae = d_sae
        self.layer_indices = layer_indices or list(range(n_layers))
        self.activation = activation
        self.sparsity_weight = sparsity_weight
        self.normalize_decoder = normalize_decoder
        self.dead_feature_threshold = dead_feature_threshold
        self.W_enc = nn.Parameter(torch.empty(n_layers, d_model, d_sae))
        self.b_enc = nn.Parameter(torch.zeros(d_sae))
        self.W_dec = nn.Parameter(torch.empty(d_sae, n_layers, d_model))
        self.b_dec = nn.Parameter(torch.zeros(n_layers, d_model))
        self._init_weights()
        self.register_buffer("feature_last_activated", torch.zeros(d_sae, dtype=torch.long))
        self.register_buffer("step_count", torch.tensor(0, dtype=torch.long))

    def _init_weights(self) -> None:
        """xavier decoder, unit rows over (L, d), x0.1; encoder = decoder^T (crosscoder.py:105-120)."""
        with torch.no_grad():
            nn.init.xavier_uniform_(self.W_dec)
            if self.normalize_decoder:
                flat = nn.functional.normalize(self.W_dec.view(self.d_sae, -1), dim=1)
                self.W_dec.data = flat.view(self.d_sae, self.n_layers, self.d_model)
                self.W_dec.data *= 0.1
            for layer in range(self.n_layers):
                self.W_enc.data[layer] = self.W_dec.data[:, layer, :].T

    def normalize_decoder_weights(self) -> None:
        with torch.no_grad():
            flat = self.W_dec.data.view(self.d_sae, -1)
            if flat.is_cuda and flat.is_contiguous():
                ops.renorm_decoder_(flat, 1e-12)
            else:
                self.W_dec.data = nn.functional.normalize(flat, dim=1).view_as(self.W_dec)

    def get_decoder_norms(self) -> Tensor:
        return torch.norm(self.W_dec.view(self.d_sae, -1), dim=1)

    def _pre_activation(self, layer_activations: dict[int, Tensor]) -> Tensor:
        first = next(iter(layer_activations.values()))
        pre = torch.zeros(first.shape[0], self.d_sae, device=first.device)
        for layer_idx, acts in layer_activations.items():
            pre = pre + torch.einsum("bd,ds->bs", acts, self.W_enc[self.layer_indices.index(layer_idx)])
        return pre + self.b_enc

    def encode(self, layer_activations: dict[int, Tensor]) -> Tensor:
        if self.activation != "relu":
            raise ValueError(f"Unknown activation: {self.activation}")
        return torch.relu(self._pre_activation(layer_activations))

    def decode(self, hidden: Tensor) -> dict[int, Tensor]:
        return {layer_idx: torch.einsum("bs,sd->bd", hidden, self.W_dec[:, i, :]) + self.b_dec[i]
                for i, layer_idx in enumerate(self.layer_indices)}

    def forward(self, layer_activations: dict[int, Tensor]) -> CrosscoderOutput:
        hidden = self.encode(layer_activations)
        reconstructed = self.decode(hidden)
        per_layer = {li: torch.mean((rec - layer_activations[li]) ** 2) for li, rec in reconstructed.items()}
        total = torch.tensor(0.0, device=hidden.device)
        for v in per_layer.values():
            total = total + v
        sparsity = torch.mean(hidden.abs() @ self.get_decoder_norms())
        loss = total + self.sparsity_weight * sparsity
        l0 = (hidden > 0).float().sum(dim=-1).mean()
        self._update_dead_features(hidden)
        return CrosscoderOutput(reconstructed, hidden, loss, total, sparsity, l0, per_layer)

    def _update_dead_features(self, hidden: Tensor) -> None:
        if self.training:
            self.step_count += 1
            self.feature_last_activated[(hidden > 0).any(dim=0)] = self.step_count

    def get_dead_features(self) -> Tensor:
        return (self.step_count - self.feature_last_activated) > self.dead_feature_threshold

    def get_dead_feature_ratio(self) -> float:
        return self.get_dead_features().float().mean().item()

    def get_feature_layer_norms(self) -> Tensor:
        return torch.norm(self.W_dec, dim=2)

    def get_cross_layer_features(self, threshold: float = 0.1) -> Tensor:
        norms = self.get_feature_layer_norms()
        rel = norms / (norms.max(dim=1, keepdim=True).values + 1e-8)
        return (rel > threshold).sum(dim=1) >= 2


class TopKCrossLayerCrosscoder(CrossLayerCrosscoder):
    """TopK crosscoder (crosscoder.py:298-379) on the fused sm_100a path."""

    def __init__(self, d_model: int, n_layers: int, d_sae: int, k: int = 32,
                 layer_indices: list[int] | None = None, normalize_decoder: bool = True,
                 dead_feature_threshold: int = 10_000, precision: str | None = None):
        super().__init__(d_model=d_model, n_layers=n_layers, d_sae=d_sae, layer_indices=layer_indices,
                         activation="relu", sparsity_weight=0.0, normalize_decoder=normalize_decoder,
                         dead_feature_threshold=dead_feature_threshold)
        self.k = k
        if precision not in (None, "bf16", "fp32"):
            raise ValueError("precision must be None, 'bf16' or 'fp32'")
        self.precision = precision
        self._global_rows: int | None = None

    def _use_bf16(self) -> bool:
        if self.precision is not None:
            return self.precision == "bf16"
        return torch.is_autocast_enabled("cuda")

    def _concat(self, layer_activations: dict[int, Tensor]) -> Tensor:
        """[B, L*d] in internal layer order; layers that were not given contribute zeros to the
        encoder sum (crosscoder.py:151-157 iterates only the given layers)."""
        first = next(iter(layer_activations.values()))
        if not first.is_cuda:
            raise RuntimeError("TopKCrossLayerCrosscoder (whisper_sae_b200) runs on CUDA sm_100a only")
        cols = []
        for li in self.layer_indices:
            a = layer_activations.get(li)
            cols.append(a.to(torch.float32) if a is not None else torch.zeros_like(first, dtype=torch.float32))
        for li in layer_activations:
            self.layer_indices.index(li)       # unknown layer => ValueError, like the reference
        return torch.cat(cols, dim=1)

    def _w_enc_cat(self) -> Tensor:
        """W_enc [L, d, F] viewed as the [F, L*d] matrix of the concatenated-input GEMM."""
        return self.W_enc.permute(2, 0, 1).reshape(self.d_sae, self.n_layers * self.d_model)

    def encode(self, layer_activations: dict[int, Tensor]) -> Tensor:
        x = self._concat(layer_activations).contiguous()
        if torch.is_grad_enabled() and (x.requires_grad or self.W_enc.requires_grad):
            _warn_encode_detached()       # no grad_fn on the result (see TopKSAE.encode)
        terms = 1 if self._use_bf16() else _fp32_terms()
        a = ops.pack_activations(x, None, terms)
        w = ops.pack_encoder(self._w_enc_cat().detach().contiguous(), self.b_enc.detach(), terms)
        idx, val = ops.encode_topk(a, w, x.shape[0], self.d_sae, x.shape[1], terms, self.k)
        return ops.densify_hidden(idx, val, self.d_sae)

    def forward(self, layer_activations: dict[int, Tensor]) -> CrosscoderOutput:
        missing = [li for li in self.layer_indices if li not in layer_activations]
        if missing:    # the decoder reconstructs every layer and the loss indexes the inputs by it
            raise KeyError(missing[0])
        x = self._concat(layer_activations)
        L, d, F = self.n_layers, self.d_model, self.d_sae
        st = _SparseState()
        w_dec = self.W_dec.view(F, L * d).t()                    # [L*d, F] view, feature-major storage
        mean_loss = _FusedTopKSAE.apply(self, st, self._use_bf16(), x, None, None, self._w_enc_cat(),
                                        self.b_enc, w_dec, self.b_dec.view(L * d))
        loss = mean_loss * float(L)     # sum_l mean_l == L * mean over the concatenation
        self._last_sparse = st
        B = x.shape[0]
        xd = x.detach()

        def _recon() -> dict[int, Tensor]:
            rec = st.resid + xd
            return {li: rec[:, i * d:(i + 1) * d] for i, li in enumerate(self.layer_indices)}

        def _per_layer() -> dict[int, Tensor]:
            r2 = (st.resid ** 2).view(B, L, d).mean(dim=(0, 2))
            return {li: r2[i] for i, li in enumerate(self.layer_indices)}

        return CrosscoderOutput(
            _recon, lambda: ops.densify_hidden(st.idx, st.val, F), loss, loss,
            lambda: torch.zeros((), dtype=torch.float32, device=x.device),
            lambda: st.stats[1].to(torch.float32) / float(B), _per_layer)


def create_crosscoder(d_model: int, n_layers: int, d_sae: int, k: int | None = None,
                      use_topk: bool = True, **kwargs) -> nn.Module:
    """Factory (crosscoder.py:382-417)."""
    if use_topk:
        return TopKCrossLayerCrosscoder(d_model=d_model, n_layers=n_layers, d_sae=d_sae, k=k or 32, **kwargs)
    return CrossLayerCrosscoder(d_model=d_model, n_layers=n_layers, d_sae=d_sae, **kwargs)
