"""Activation cache (drop-in for the hot-path part of ``whisper_sae.data``)."""

from .feature_cache import CacheMetadata, FeatureCache, ResidentBatches, extract_and_cache_features

__all__ = ["CacheMetadata", "FeatureCache", "ResidentBatches", "extract_and_cache_features"]
