"""SAE modules and trainer (drop-in for ``whisper_sae.sae``)."""

from .model import ReLUSAE, SAEOutput, TopKSAE, create_sae
from .training import SAETrainer, TrainingMetrics

__all__ = ["ReLUSAE", "SAEOutput", "TopKSAE", "create_sae", "SAETrainer", "TrainingMetrics"]
