"""Typed experiment configuration (Pydantic v2), schema-compatible with the reference's
``whisper_sae.config`` (/root/reference/src/whisper_sae/config.py:10-177) so the same YAML
files (``configs/tiny_test.yaml`` / ``tiny_default.yaml``) load unchanged and ``model_dump()``
round-trips through trainer checkpoints.

Only the *schema* is shared with the reference — names, defaults and bounds — and it is written
down here once, as data (``_SCHEMA``); the model classes are generated from that table with
``pydantic.create_model``.  No kernel knob is added here: tuning goes through constructor kwargs /
environment variables of the kernels.
"""

from __future__ import annotations

from pathlib import Path
from typing import Any, Literal

import yaml
from pydantic import BaseModel, Field, create_model

# (d_model, encoder layers, decoder layers) per published Whisper checkpoint; config.py:25-33
_WHISPER_SHAPES: dict[str, tuple[int, int, int]] = {
    f"openai/whisper-{name}": shape
    for name, shape in {
        "tiny": (384, 4, 4), "base": (512, 6, 6), "small": (768, 12, 12), "medium": (1024, 24, 24),
        "large": (1280, 32, 32), "large-v2": (1280, 32, 32), "large-v3": (1280, 32, 32),
    }.items()
}


def _layers_0_to_3() -> list[int]:
    return [0, 1, 2, 3]


# section -> [(field, type, default, constraints)]   (reference lines in the trailing comments)
_SCHEMA: dict[str, list[tuple[str, Any, Any, dict]]] = {
    "Whisper": [                                                        # config.py:10-37
        ("model_name", str, "openai/whisper-tiny", {}),
        ("hidden_dim", int, 384, {}),
        ("num_encoder_layers", int, 4, {}),
        ("num_decoder_layers", int, 4, {}),
    ],
    "SAE": [                                                            # config.py:40-75
        ("expansion_factor", int, 8, {"ge": 4, "le": 32}),
        ("activation", Literal["topk", "relu", "gelu"], "topk", {}),
        ("k", int, 32, {"ge": 1}),
        ("normalize_decoder", bool, True, {}),
        ("dead_feature_threshold", int, 10_000, {}),
        ("dead_feature_resample", bool, True, {}),
    ],
    "Training": [                                                       # config.py:78-93
        ("batch_size", int, 128, {"ge": 1}),
        ("learning_rate", float, 1e-4, {"gt": 0}),
        ("weight_decay", float, 0.0, {"ge": 0}),
        ("epochs", int, 50, {"ge": 1}),
        ("warmup_steps", int, 1000, {"ge": 0}),
        ("gradient_clip", float, 1.0, {"gt": 0}),
        ("use_amp", bool, True, {}),
        ("checkpoint_every", int, 10, {}),
        ("seed", int, 42, {}),
        ("num_workers", int, 4, {"ge": 0}),
    ],
    "Data": [                                                           # config.py:96-108
        ("dataset_name", str, "librispeech_asr", {}),
        ("dataset_subset", str, "clean", {}),
        ("dataset_split", str, "train.100", {}),
        ("max_samples", int, 100_000, {"ge": 1}),
        ("cache_dir", Path, Path("cache"), {}),
        ("streaming", bool, True, {}),
    ],
    "Wandb": [                                                          # config.py:111-122
        ("enabled", bool, True, {}),
        ("project", str, "whisper-sae", {}),
        ("entity", str | None, None, {}),
        ("name", str | None, None, {}),
        ("tags", list[str], list, {}),
        ("log_every", int, 100, {}),
    ],
}


def _model(section: str, base: type[BaseModel] = BaseModel, extra: dict | None = None) -> type[BaseModel]:
    fields: dict[str, Any] = {}
    for name, tp, default, bounds in _SCHEMA.get(section, []):
        fields[name] = (tp, Field(default_factory=default, **bounds) if callable(default)
                        else Field(default=default, **bounds))
    fields.update(extra or {})
    return create_model(f"{section}Config", __base__=base, __module__=__name__, **fields)


class _WhisperBase(BaseModel):
    def model_post_init(self, __context: Any) -> None:
        """Known checkpoints override the three shape fields (config.py:25-37)."""
        shape = _WHISPER_SHAPES.get(self.model_name)
        if shape is not None:
            self.hidden_dim, self.num_encoder_layers, self.num_decoder_layers = shape


class _SAEBase(BaseModel):
    def get_hidden_dim(self, input_dim: int) -> int:
        """SAE width = input width x expansion factor (config.py:73-75)."""
        return self.expansion_factor * input_dim


WhisperConfig = _model("Whisper", _WhisperBase)
SAEConfig = _model("SAE", _SAEBase)
TrainingConfig = _model("Training")
DataConfig = _model("Data")
WandbConfig = _model("Wandb")


class _ExperimentBase(BaseModel):
    """Root of the tree + YAML round trip (config.py:125-160)."""

    @classmethod
    def from_yaml(cls, path: str | Path):
        with open(path) as fh:
            return cls(**(yaml.safe_load(fh) or {}))

    def to_yaml(self, path: str | Path) -> None:
        with open(path, "w") as fh:
            yaml.dump(self.model_dump(mode="json"), fh, default_flow_style=False)

    def get_run_dir(self) -> Path:
        out = Path(self.output_dir) / self.experiment_name
        out.mkdir(parents=True, exist_ok=True)
        return out


ExperimentConfig = _model("Experiment", _ExperimentBase, {
    "whisper": (WhisperConfig, Field(default_factory=WhisperConfig)),
    "sae": (SAEConfig, Field(default_factory=SAEConfig)),
    "training": (TrainingConfig, Field(default_factory=TrainingConfig)),
    "data": (DataConfig, Field(default_factory=DataConfig)),
    "wandb": (WandbConfig, Field(default_factory=WandbConfig)),
    "encoder_layers": (list[int], Field(default_factory=_layers_0_to_3)),
    "decoder_layers": (list[int], Field(default_factory=_layers_0_to_3)),
    "output_dir": (Path, Field(default=Path("outputs"))),
    "experiment_name": (str, Field(default="default")),
})


class _LayerBase(BaseModel):
    """One (component, layer) training unit (config.py:163-177; used by tests only upstream)."""

    @property
    def name(self) -> str:
        return "_".join((self.component, f"layer{self.layer_idx}"))

    @property
    def hidden_dim(self) -> int:
        return self.sae_config.get_hidden_dim(self.input_dim)


LayerConfig = _model("Layer", _LayerBase, {
    "component": (Literal["encoder", "decoder"], ...),
    "layer_idx": (int, Field(ge=0)),
    "input_dim": (int, ...),
    "sae_config": (SAEConfig, Field(default_factory=SAEConfig)),
    "training_config": (TrainingConfig, Field(default_factory=TrainingConfig)),
})
