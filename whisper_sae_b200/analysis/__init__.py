"""Feature-interpretation tools that consume the hot path's sparse code (analysis/__init__.py of
the reference, minus audio-clip extraction, which needs the audio files)."""

from .feature_viz import (
    FeatureActivation,
    FeatureInterpretation,
    FeatureReport,
    TopKTracker,
    collect_top_activations,
)

__all__ = ["FeatureActivation", "FeatureInterpretation", "FeatureReport", "TopKTracker",
           "collect_top_activations"]
