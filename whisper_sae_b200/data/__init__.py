"""Activation cache (drop-in for the hot-path part of ``whisper_sae.data``)."""

from .feature_cache import CacheMetadata, FeatureCache, ResidentBatches

__all__ = ["CacheMetadata", "FeatureCache", "ResidentBatches"]
