"""Typed experiment configuration (Pydantic v2), field-compatible with the reference's
``whisper_sae.config`` (/root/reference/src/whisper_sae/config.py:10-177) so the same YAML
files (``configs/tiny_test.yaml`` / ``tiny_default.yaml``) load unchanged.

Only the *schema* is shared with the reference: names, defaults and bounds.  No kernel knob is
added here — tuning goes through constructor kwargs / environment variables of the kernels.
"""

from __future__ import annotations

from pathlib import Path
from typing import Literal

import yaml
from pydantic import BaseModel, Field, model_validator

# (d_model, encoder layers, decoder layers) per published Whisper checkpoint; config.py:25-33
_WHISPER_SHAPES: dict[str, tuple[int, int, int]] = {
    "openai/whisper-tiny": (384, 4, 4),
    "openai/whisper-base": (512, 6, 6),
    "openai/whisper-small": (768, 12, 12),
    "openai/whisper-medium": (1024, 24, 24),
    "openai/whisper-large": (1280, 32, 32),
    "openai/whisper-large-v2": (1280, 32, 32),
    "openai/whisper-large-v3": (1280, 32, 32),
}


class WhisperConfig(BaseModel):
    model_name: str = Field(default="openai/whisper-tiny")
    hidden_dim: int = Field(default=384)
    num_encoder_layers: int = Field(default=4)
    num_decoder_layers: int = Field(default=4)

    @model_validator(mode="after")
    def _fill_from_model_name(self) -> "WhisperConfig":
        shape = _WHISPER_SHAPES.get(self.model_name)
        if shape is not None:
            self.hidden_dim, self.num_encoder_layers, self.num_decoder_layers = shape
        return self


class SAEConfig(BaseModel):
    expansion_factor: int = Field(default=8, ge=4, le=32)
    activation: Literal["topk", "relu", "gelu"] = Field(default="topk")
    k: int = Field(default=32, ge=1)
    normalize_decoder: bool = Field(default=True)
    dead_feature_threshold: int = Field(default=10_000)
    dead_feature_resample: bool = Field(default=True)

    def get_hidden_dim(self, input_dim: int) -> int:
        return input_dim * self.expansion_factor


class TrainingConfig(BaseModel):
    batch_size: int = Field(default=128, ge=1)
    learning_rate: float = Field(default=1e-4, gt=0)
    weight_decay: float = Field(default=0.0, ge=0)
    epochs: int = Field(default=50, ge=1)
    warmup_steps: int = Field(default=1000, ge=0)
    gradient_clip: float = Field(default=1.0, gt=0)
    use_amp: bool = Field(default=True)
    checkpoint_every: int = Field(default=10)
    seed: int = Field(default=42)
    num_workers: int = Field(default=4, ge=0)


class DataConfig(BaseModel):
    dataset_name: str = Field(default="librispeech_asr")
    dataset_subset: str = Field(default="clean")
    dataset_split: str = Field(default="train.100")
    max_samples: int = Field(default=100_000, ge=1)
    cache_dir: Path = Field(default=Path("cache"))
    streaming: bool = Field(default=True)


class WandbConfig(BaseModel):
    enabled: bool = Field(default=True)
    project: str = Field(default="whisper-sae")
    entity: str | None = Field(default=None)
    name: str | None = Field(default=None)
    tags: list[str] = Field(default_factory=list)
    log_every: int = Field(default=100)


class ExperimentConfig(BaseModel):
    whisper: WhisperConfig = Field(default_factory=WhisperConfig)
    sae: SAEConfig = Field(default_factory=SAEConfig)
    training: TrainingConfig = Field(default_factory=TrainingConfig)
    data: DataConfig = Field(default_factory=DataConfig)
    wandb: WandbConfig = Field(default_factory=WandbConfig)
    encoder_layers: list[int] = Field(default_factory=lambda: [0, 1, 2, 3])
    decoder_layers: list[int] = Field(default_factory=lambda: [0, 1, 2, 3])
    output_dir: Path = Field(default=Path("outputs"))
    experiment_name: str = Field(default="default")

    @classmethod
    def from_yaml(cls, path: str | Path) -> "ExperimentConfig":
        return cls(**yaml.safe_load(Path(path).read_text()))

    def to_yaml(self, path: str | Path) -> None:
        Path(path).write_text(yaml.dump(self.model_dump(mode="json"), default_flow_style=False))

    def get_run_dir(self) -> Path:
        run_dir = self.output_dir / self.experiment_name
        run_dir.mkdir(parents=True, exist_ok=True)
        return run_dir


class LayerConfig(BaseModel):
    component: Literal["encoder", "decoder"]
    layer_idx: int = Field(ge=0)
    input_dim: int
    sae_config: SAEConfig = Field(default_factory=SAEConfig)
    training_config: TrainingConfig = Field(default_factory=TrainingConfig)

    @property
    def name(self) -> str:
        return f"{self.component}_layer{self.layer_idx}"

    @property
    def hidden_dim(self) -> int:
        return self.sae_config.get_hidden_dim(self.input_dim)
