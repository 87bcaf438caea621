"""whisper_sae_b200 — B200-native (sm_100a) implementation of whisper-sae's TopK-SAE train step.

Public surface mirrors the reference package ``whisper_sae`` for the hot path only:
``config`` (Pydantic models), ``sae`` (TopKSAE / ReLUSAE / create_sae / SAETrainer) and
``data`` (FeatureCache).  The compute lives in ``csrc/`` behind the C ABI in ``include/wsae.h``.
"""

__version__ = "0.1.0"

from .config import (DataConfig, ExperimentConfig, LayerConfig, SAEConfig, TrainingConfig,
                     WandbConfig, WhisperConfig)

__all__ = ["DataConfig", "ExperimentConfig", "LayerConfig", "SAEConfig", "TrainingConfig",
           "WandbConfig", "WhisperConfig", "__version__"]
