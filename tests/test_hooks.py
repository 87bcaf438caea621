"""Activation extraction (reference sae/hooks.py, data/feature_cache.py:200-306).

CPU: the oracle's LayerNorm + flatten against activations extracted by the live reference from a
seeded random-init Whisper (oracle/make_golden_hooks.py -> tests/golden/hooks.pt).
GPU: wsae_layernorm_rows against torch.nn.functional.layer_norm and the oracle; the extractor and
extract_and_cache_features against the same fixture (the model is rebuilt from the recipe's seeds).
Tolerance: fp32 LayerNorm, 2e-5 absolute on O(1) outputs (different summation order).
"""

from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import hooks_oracle as HO

GOLDEN = Path(__file__).parent / "golden" / "hooks.pt"
TOL = 2e-5


def _fx():
    return torch.load(GOLDEN, weights_only=False)


def _build(recipe):
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    cfg = WhisperConfig(**{k: v for k, v in recipe.items() if k not in ("model_seed", "input_seed", "batch")})
    torch.manual_seed(recipe["model_seed"])
    model = WhisperForConditionalGeneration(cfg).eval()
    with torch.no_grad():
        g = torch.Generator().manual_seed(recipe["model_seed"] + 1)
        for ln in (model.model.encoder.layer_norm, model.model.decoder.layer_norm):
            ln.weight.copy_(1.0 + 0.1 * torch.randn(ln.weight.shape, generator=g))
            ln.bias.copy_(0.1 * torch.randn(ln.bias.shape, generator=g))
    x = torch.randn(recipe["batch"], recipe["num_mel_bins"], 2 * recipe["max_source_positions"],
                    generator=torch.Generator().manual_seed(recipe["input_seed"]))
    return model, x


# ---------------------------------------------------------------- CPU: oracle pinned to the reference
def test_oracle_layernorm_matches_reference_extraction():
    fx = _fx()
    g, b, eps = fx["encoder_ln"]
    for layer in fx["encoder_layers"]:
        got = HO.layer_norm_rows(fx["raw_encoder"][layer].numpy(), g.numpy(), b.numpy(), eps)
        np.testing.assert_allclose(got, fx["reference_encoder"][layer].numpy(), atol=TOL, rtol=0)
        np.testing.assert_allclose(HO.flatten(got), fx["reference_encoder_flat"][layer].numpy(), atol=TOL, rtol=0)
    g, b, eps = fx["decoder_ln"]
    for layer, want in fx["decoder_expected"].items():
        got = HO.layer_norm_rows(fx["raw_decoder"][layer].numpy(), g.numpy(), b.numpy(), eps)
        np.testing.assert_allclose(got, want.numpy(), atol=TOL, rtol=0)


# ---------------------------------------------------------------- GPU: kernel
@pytest.mark.gpu
@pytest.mark.parametrize("rows,d", [(1, 64), (77, 100), (1500, 384), (513, 1280), (9, 4096), (300, 8)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_layernorm_rows_matches_torch(rows, d, dtype):
    from whisper_sae_b200 import ops
    g = torch.Generator().manual_seed(rows * d)
    x = (torch.randn(rows, d, generator=g) * 3 + 0.5).to(dtype).cuda()
    gamma = (1 + 0.2 * torch.randn(d, generator=g)).cuda()
    beta = (0.3 * torch.randn(d, generator=g)).cuda()
    out = torch.full((rows + 5, d), float("nan"), device="cuda")
    ops.layernorm_rows_(x, gamma, beta, 1e-5, out, row0=3)
    want = torch.nn.functional.layer_norm(x.float(), (d,), gamma, beta, 1e-5)
    torch.testing.assert_close(out[3:3 + rows], want, atol=TOL, rtol=1e-5)
    assert torch.isnan(out[:3]).all() and torch.isnan(out[3 + rows:]).all()     # neighbours untouched
    np.testing.assert_allclose(out[3:3 + rows].cpu().numpy(),
                               HO.layer_norm_rows(x.float().cpu().numpy(), gamma.cpu().numpy(),
                                                  beta.cpu().numpy(), 1e-5), atol=TOL, rtol=1e-5)


@pytest.mark.gpu
def test_layernorm_rows_rejects_bad_arguments():
    from whisper_sae_b200 import ops
    x = torch.zeros(4, 16, device="cuda")
    with pytest.raises(RuntimeError):
        ops.layernorm_rows_(x, None, None, 1e-5, torch.zeros(3, 16, device="cuda"))          # does not fit
    with pytest.raises(RuntimeError):
        ops.layernorm_rows_(x.cpu(), None, None, 1e-5, torch.zeros(4, 16, device="cuda"))    # CPU input
    with pytest.raises(RuntimeError):
        ops.layernorm_rows_(torch.zeros(2, 8192, device="cuda"), None, None, 1e-5,
                            torch.zeros(2, 8192, device="cuda"))                             # d > 4096


# ---------------------------------------------------------------- GPU: extractor vs the reference's output
@pytest.mark.gpu
def test_extract_features_batch_matches_reference_golden():
    from whisper_sae_b200.sae.hooks import extract_features_batch, flatten_activations
    fx = _fx()
    model, x = _build(fx["recipe"])
    model = model.cuda()
    dec_layers = sorted(fx["decoder_expected"])
    got = extract_features_batch(model, x, encoder_layers=fx["encoder_layers"], decoder_layers=dec_layers,
                                 device="cuda")
    for layer in fx["encoder_layers"]:
        a = got["encoder"][layer]
        assert a.is_cuda and a.dtype == torch.float32
        # the GPU forward itself differs from the CPU one at the 1e-5 level (TF32 off, different GEMM order)
        torch.testing.assert_close(a.cpu(), fx["reference_encoder"][layer], atol=2e-4, rtol=1e-4)
        assert flatten_activations(a, "encoder").shape == fx["reference_encoder_flat"][layer].shape
    for layer in dec_layers:      # the reference's intent for the decoder (its hook is broken here)
        torch.testing.assert_close(got["decoder"][layer].cpu(), fx["decoder_expected"][layer], atol=2e-4, rtol=1e-4)
    # without the final LayerNorm the raw layer outputs come back
    raw = extract_features_batch(model, x, fx["encoder_layers"], [], apply_layer_norm=False, device="cuda")
    torch.testing.assert_close(raw["encoder"][0].cpu(), fx["raw_encoder"][0], atol=2e-4, rtol=1e-4)


@pytest.mark.gpu
def test_extract_and_cache_features_writes_the_reference_cache_format(tmp_path):
    from whisper_sae_b200.config import DataConfig, WhisperConfig as WCfg
    from whisper_sae_b200.data import FeatureCache, extract_and_cache_features
    fx = _fx()
    model, x = _build(fx["recipe"])
    cache = FeatureCache(tmp_path, WCfg(), DataConfig())
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x), batch_size=2)   # 2 + 1 samples
    extract_and_cache_features(model, None, loader, cache, encoder_layers=[0, 2], decoder_layers=[1],
                               device="cuda")
    for layer in (0, 2):
        feats, meta = cache.load("encoder", layer)
        assert feats.dtype == torch.float32 and feats.device.type == "cpu"
        torch.testing.assert_close(feats, fx["reference_encoder_flat"][layer], atol=2e-4, rtol=1e-4)
        assert (meta.num_tokens, meta.hidden_dim, meta.num_samples) == (feats.shape[0], 64, 3)
    feats, meta = cache.load("decoder", 1)
    torch.testing.assert_close(feats, fx["decoder_expected"][1].reshape(-1, 64), atol=2e-4, rtol=1e-4)
    # the cache feeds the trainer's loader unchanged
    batch = next(iter(cache.get_dataloader("encoder", 0, batch_size=16, shuffle=False)))[0]
    assert batch.shape == (16, 64)
