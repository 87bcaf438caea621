"""CPU-side checks: config schema, C-ABI library exports, module construction / state_dict
layout, lazy SAEOutput, FeatureCache formats, launch heuristics.  No GPU compute."""

import ctypes
import json
import re
from pathlib import Path

import pytest
import torch

from oracle import topk_sae_oracle as O

ROOT = Path(__file__).resolve().parents[1]


# ------------------------------------------------------------------ config (reference tests/test_config.py)
def test_config_defaults_and_validation(tmp_path):
    from pydantic import ValidationError

    from whisper_sae_b200.config import ExperimentConfig, LayerConfig, SAEConfig, TrainingConfig, WhisperConfig

    assert WhisperConfig().hidden_dim == 384
    w = WhisperConfig(model_name="openai/whisper-large-v3")
    assert (w.hidden_dim, w.num_encoder_layers, w.num_decoder_layers) == (1280, 32, 32)
    assert WhisperConfig(model_name="custom/x", hidden_dim=99).hidden_dim == 99
    s = SAEConfig()
    assert (s.expansion_factor, s.activation, s.k, s.dead_feature_threshold) == (8, "topk", 32, 10_000)
    assert s.get_hidden_dim(384) == 3072 and s.get_hidden_dim(768) == 6144
    for bad in (dict(expansion_factor=3), dict(expansion_factor=33), dict(k=0), dict(activation="tanh")):
        with pytest.raises(ValidationError):
            SAEConfig(**bad)
    t = TrainingConfig()
    assert (t.batch_size, t.learning_rate, t.epochs, t.warmup_steps, t.gradient_clip, t.seed) == \
        (128, 1e-4, 50, 1000, 1.0, 42)
    with pytest.raises(ValidationError):
        TrainingConfig(batch_size=0)
    cfg = ExperimentConfig(experiment_name="rt", output_dir=tmp_path)
    cfg.to_yaml(tmp_path / "c.yaml")
    back = ExperimentConfig.from_yaml(tmp_path / "c.yaml")
    assert back.model_dump() == cfg.model_dump()
    assert cfg.get_run_dir() == tmp_path / "rt" and (tmp_path / "rt").is_dir()
    lc = LayerConfig(component="encoder", layer_idx=2, input_dim=384)
    assert lc.name == "encoder_layer2" and lc.hidden_dim == 3072


@pytest.mark.parametrize("fname,batch,thr,enc,dec", [("tiny_test.yaml", 64, 1000, [0], []),
                                                     ("tiny_default.yaml", 128, 10000, [0, 1, 2, 3], [0, 1, 2, 3])])
def test_shipped_yaml_configs(fname, batch, thr, enc, dec):
    from whisper_sae_b200.config import ExperimentConfig

    cfg = ExperimentConfig.from_yaml(ROOT / "configs" / fname)
    assert cfg.training.batch_size == batch and cfg.sae.dead_feature_threshold == thr
    assert cfg.encoder_layers == enc and cfg.decoder_layers == dec
    assert cfg.sae.k == 32 and cfg.sae.expansion_factor == 8 and cfg.whisper.hidden_dim == 384


# ------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    from whisper_sae_b200 import _lib

    header = (ROOT / "include" / "wsae.h").read_text()
    declared = set(re.findall(r"^int\s+(wsae_\w+)\s*\(", header, flags=re.M))
    assert declared, "no declarations parsed from include/wsae.h"
    lib = _lib.load()                       # builds with nvcc if the .so is absent; no GPU needed
    assert _lib.LIB_PATH.exists()
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for sym in sorted(declared):
        assert hasattr(raw, sym), f"{sym} declared in wsae.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), "ctypes signature table out of sync with wsae.h"
    assert lib.wsae_abi_version() == 111
    # pure host helpers can be called without a GPU
    dp, used, kp = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.wsae_packed_k(384, 1, ctypes.byref(dp), ctypes.byref(used), ctypes.byref(kp)) == 0
    assert (dp.value, used.value, kp.value) == (384, 400, 448)
    assert lib.wsae_packed_k(384, 2, None, None, None) == -1
    assert lib.wsae_encode_effective_splits(3072, 5) == 4      # 12 tiles / ceil(12/5)=3 per split
    assert lib.wsae_encode_effective_splits(100, 7) == 1


def test_sass_contains_blackwell_instructions():
    import shutil
    import subprocess

    from whisper_sae_b200 import _lib

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    _lib.load()
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing: the encoder GEMM is not on tcgen05/TMA"
    # the dense GEMMs (K1 encoder, K4 weight gradients) must be tcgen05 only - no legacy warp-level MMA.
    # (K23 may use HMMA.16816 for its 32 gathered-row dot products per activation row: operands already in
    # registers, 2 % of a GEMM tile, kernel bound by the L2 gather - csrc/wsae_decode_backward.cu.)
    lib_dir = _lib.LIB_PATH.parent
    for obj in ("wsae_encode_topk.o", "wsae_wgrad_gemm.o"):
        if not (lib_dir / obj).exists():
            continue
        o = subprocess.run(["cuobjdump", "-sass", str(lib_dir / obj)], capture_output=True, text=True).stdout
        assert "UTCHMMA" in o and "HMMA." not in o.replace("UTCHMMA", ""), f"{obj}: legacy HMMA in a dense GEMM"


# ------------------------------------------------------------------ module surface
def test_topk_sae_construction_matches_reference_layout():
    from whisper_sae_b200.sae import TopKSAE

    torch.manual_seed(42)
    sae = TopKSAE(384, 3072, k=32)
    torch.manual_seed(42)
    ref = O.init_state(384, 3072)           # pinned to the reference RNG sequence by the golden tests
    sd = sae.state_dict()
    assert list(sd) == list(ref)
    for n, t in sd.items():
        assert t.dtype == ref[n].dtype and tuple(t.shape) == tuple(ref[n].shape)
        assert torch.equal(t.contiguous(), ref[n]), n
    assert [n for n, _ in sae.named_parameters()] == list(O.PARAM_ORDER)
    torch.testing.assert_close(sae.decoder.weight.norm(dim=0), torch.full((3072,), 0.1), atol=1e-5, rtol=0)
    # feature-major storage behind the [d, F] view
    assert sae.decoder.weight.shape == (384, 3072) and sae.decoder.weight.stride() == (1, 384)
    assert sae.encoder.in_features == 384 and sae.encoder.out_features == 3072
    sae.normalize_decoder_weights()
    torch.testing.assert_close(sae.decoder.weight.norm(dim=0), torch.ones(3072), atol=1e-5, rtol=0)
    assert not bool(sae.get_dead_features().any()) and sae.get_dead_feature_ratio() == 0.0


def test_no_cpu_fallback():
    from whisper_sae_b200 import ops
    from whisper_sae_b200.sae import TopKSAE

    sae = TopKSAE(32, 64, k=4)
    with pytest.raises(RuntimeError, match="CUDA"):
        sae(torch.randn(2, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        sae.encode(torch.randn(2, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.renorm_decoder_(torch.randn(8, 4))


def test_reference_state_dict_loads(tmp_path):
    """A checkpoint with the reference layout (contiguous [d, F] decoder) loads and keeps our layout."""
    from whisper_sae_b200.sae import TopKSAE

    torch.manual_seed(3)
    ref_state = O.init_state(64, 256)
    torch.save(ref_state, tmp_path / "sae_final.pt")
    sae = TopKSAE(64, 256, k=8)
    sae.load_state_dict(torch.load(tmp_path / "sae_final.pt"))
    assert sae.decoder.weight.stride() == (1, 64)
    assert torch.equal(sae.decoder.weight.contiguous(), ref_state["decoder.weight"])
    out = tmp_path / "again.pt"
    torch.save(sae.state_dict(), out)
    again = torch.load(out)
    assert list(again) == list(ref_state) and torch.equal(again["decoder.weight"].contiguous(),
                                                          ref_state["decoder.weight"])


def test_create_sae_factory_and_relu_sae():
    from whisper_sae_b200.config import SAEConfig
    from whisper_sae_b200.sae import ReLUSAE, SAEOutput, TopKSAE, create_sae

    m = create_sae(SAEConfig(expansion_factor=8, activation="topk", k=32), input_dim=384)
    assert isinstance(m, TopKSAE) and m.hidden_dim == 3072 and m.k == 32
    for act in ("relu", "gelu"):
        r = create_sae(SAEConfig(activation=act), input_dim=512)
        assert isinstance(r, ReLUSAE) and r.hidden_dim == 4096
    relu = ReLUSAE(32, 64, sparsity_weight=0.01)
    out = relu(torch.randn(5, 32))             # dense ReLU SAE is plain torch: works on CPU
    assert isinstance(out, SAEOutput)
    assert out.loss.item() == pytest.approx(
        (out.reconstruction_loss + 0.01 * out.sparsity_loss).item(), rel=1e-6)
    torch.testing.assert_close(relu.decoder.weight.norm(dim=0), torch.ones(64), atol=1e-5, rtol=0)


def test_sae_output_is_lazy_and_tuple_like():
    from whisper_sae_b200.sae import SAEOutput

    calls = []

    def make(tag, t):
        def thunk():
            calls.append(tag)
            return t
        return thunk

    a, b = torch.ones(2, 3), torch.zeros(2, 5)
    s = torch.tensor(1.5)
    out = SAEOutput(make("r", a), make("h", b), s, s, torch.tensor(0.0), torch.tensor(2.0))
    assert calls == [] and out.loss is out.reconstruction_loss
    assert out.hidden is b and out.hidden is b and calls == ["h"]
    assert len(out) == 6 and out[0] is a and out[-1].item() == 2.0
    r, h, loss, rl, sl, l0 = out
    assert r is a and calls == ["h", "r"]
    assert set(out._asdict()) == set(SAEOutput._fields)


# ------------------------------------------------------------------ trainer (host side only)
def test_trainer_host_state(tmp_path):
    from whisper_sae_b200.config import TrainingConfig
    from whisper_sae_b200.sae import SAETrainer, TopKSAE

    cfg = TrainingConfig(batch_size=16, epochs=2, warmup_steps=10, use_amp=True)
    tr = SAETrainer(TopKSAE(64, 128, k=8), cfg, device="cpu", run_dir=tmp_path / "r")
    assert (tmp_path / "r").is_dir()
    assert tr.use_amp is False and not tr.scaler.is_enabled()      # AMP is CUDA-only (training.py:73-75)
    assert (tr.global_step, tr.epoch, tr.metrics_history, tr.num_resampled_total) == (0, 0, [], 0)
    assert tr.scheduler is None and tr.wandb_run is None and tr._resample_dataset is None
    tr.setup_scheduler(200)
    lrs = []
    for _ in range(30):
        lrs.append(tr.optimizer.param_groups[0]["lr"])
        tr.optimizer.step()
        tr.scheduler.step()
    want = O.lr_sequence(cfg.learning_rate, 200, 10)[:30]
    assert lrs == pytest.approx(want, rel=1e-9)
    assert lrs[0] == pytest.approx(0.01 * cfg.learning_rate)
    ck = torch.load(tr.save_checkpoint("c.pt"), weights_only=False)
    assert list(ck["model_state_dict"])[0] == "b_pre" and ck["config"]["batch_size"] == 16
    ds = torch.utils.data.TensorDataset(torch.randn(10, 64))
    tr.set_resample_dataset(ds)
    assert tr._resample_dataset is ds and tr._maybe_resample_dead_features() == 0


# ------------------------------------------------------------------ feature cache
def test_feature_cache_formats(tmp_path):
    from whisper_sae_b200.config import DataConfig, WhisperConfig
    from whisper_sae_b200.data import CacheMetadata, FeatureCache, ResidentBatches

    fc = FeatureCache(tmp_path / "features", WhisperConfig(), DataConfig(max_samples=7))
    assert not fc.has_cache("encoder", 0)
    feats = O.synthetic_activations(100, 384, seed=1)
    fc.save(feats, "encoder", 0, num_samples=7)
    assert (tmp_path / "features" / "whisper-tiny_encoder_layer0.pt").exists()
    meta = json.loads((tmp_path / "features" / "whisper-tiny_encoder_layer0_meta.json").read_text())
    assert set(meta) == {"model_name", "component", "layer_idx", "hidden_dim", "num_samples",
                         "num_tokens", "created_at", "data_config"}
    assert meta["hidden_dim"] == 384 and meta["num_tokens"] == 100 and meta["data_config"]["cache_dir"] == "cache"
    assert fc.has_cache("encoder", 0) and not fc.has_cache("decoder", 0)
    loaded, m = fc.load("encoder", 0)
    assert isinstance(m, CacheMetadata) and torch.equal(loaded, feats)
    dl = fc.get_dataloader("encoder", 0, batch_size=32, shuffle=False)
    batches = list(dl)
    assert len(batches) == 4 and isinstance(batches[0], list) and batches[0][0].shape == (32, 384)
    rb = fc.get_dataloader("encoder", 0, batch_size=32, shuffle=True, device="cpu")
    assert isinstance(rb, ResidentBatches) and len(rb) == 4
    seen = torch.cat([b[0] for b in rb])
    assert seen.shape == feats.shape
    assert torch.equal(torch.sort(seen[:, 0]).values, torch.sort(feats[:, 0]).values)   # a permutation


# ------------------------------------------------------------------ launch heuristics
def test_choose_nsplit():
    from whisper_sae_b200.ops import choose_nsplit

    assert choose_nsplit(64, 3072, 32, 148) == 12          # one row block: spread F over 12 CTAs
    assert choose_nsplit(128 * 148, 3072, 32, 148) == 1    # exactly one wave already
    assert choose_nsplit(65536, 3072, 32, 148) == 1        # per-item start-up cost beats tail filling
    assert choose_nsplit(8, 128, 4, 148) == 1              # a single tile cannot be split
    assert choose_nsplit(64, 40960, 64, 148) <= 32         # merge kernel limit: nsplit*k <= 2048


def test_row_step_shape_rule(monkeypatch):
    """Which batches take the one-block-per-row small-batch step (wsae_row_step)."""
    from whisper_sae_b200.ops import row_step_supported

    assert row_step_supported(128, 384, 3072, 32, True)             # the shipped YAML batch
    assert row_step_supported(512, 768, 6144, 32, True)
    assert not row_step_supported(1024, 384, 3072, 32, True)        # large-batch chain (K23 + K4) from here on
    assert not row_step_supported(128, 384, 3072, 64, True)         # k > 32
    assert not row_step_supported(128, 384, 3072, 32, False)        # fp32-grade mode keeps K2 + K3
    assert not row_step_supported(128, 1280, 65536, 32, True)       # keys + row do not fit shared memory
    monkeypatch.setenv("WSAE_ROW_STEP_ROWS", "0")
    assert not row_step_supported(128, 384, 3072, 32, True)


def test_graphed_variant_step_refuses_cpu_and_plain_optimizers():
    from whisper_sae_b200.sae import GraphedVariantStep, make_optimizer
    from whisper_sae_b200.sae.transcoder import TopKTranscoder

    m = TopKTranscoder(16, 16, 32, k=4)
    opt = make_optimizer(m, lr=1e-3)
    assert all(g["capturable"] for g in opt.param_groups)
    with pytest.raises(RuntimeError):
        GraphedVariantStep(m, opt)                                   # CPU module: no fallback


def test_process_group_teardown_guard_releases_graphs_first(monkeypatch):
    """WSAE_DP_GRAPH=1 mode: live trainers drop their captured graphs before the communicator goes away."""
    import torch.distributed as dist

    from whisper_sae_b200.sae import training

    calls = []

    class FakeTrainer:
        def release_graphs(self):
            calls.append("release")

    monkeypatch.setattr(dist, "destroy_process_group", lambda *a, **k: calls.append("destroy"))
    monkeypatch.setattr(training, "_DP_GRAPH_TRAINERS", __import__("weakref").WeakSet())
    t = FakeTrainer()
    training._guard_process_group_teardown(t)
    training._guard_process_group_teardown(t)                        # installs the wrapper once
    dist.destroy_process_group()
    assert calls == ["release", "destroy"]
