"""CPU restatement of the reference's activation post-processing — TEST INFRASTRUCTURE ONLY.

Only ``tests/`` may import this module; the product path (``whisper_sae_b200/sae/hooks.py`` ->
``wsae_layernorm_rows``) never does.

Follows /root/reference/src/whisper_sae/sae/hooks.py:
  :85-86, :103-104  the model's final LayerNorm applied to every hooked hidden state
                    (torch.nn.LayerNorm: biased variance, eps inside the square root)
  :213-230          flatten_activations [batch, seq, d] -> [batch * seq, d]
Pinned by tests/golden/hooks.pt: activations extracted by the live reference from a seeded
random-init Whisper (oracle/make_golden_hooks.py).
"""

from __future__ import annotations

import numpy as np


def layer_norm_rows(hidden, gamma, beta, eps: float) -> np.ndarray:
    x = np.asarray(hidden, dtype=np.float64)
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
    y = (x - mean) / np.sqrt(var + eps)
    return (y * np.asarray(gamma, dtype=np.float64) + np.asarray(beta, dtype=np.float64)).astype(np.float32)


def flatten(acts) -> np.ndarray:
    a = np.asarray(acts)
    return a.reshape(-1, a.shape[-1])
