"""CPU restatement of the reference's per-feature top-k tracker — TEST INFRASTRUCTURE ONLY.

Only ``tests/`` and ``bench.py``'s CPU-baseline legs may import this module; the product path
(``whisper_sae_b200/analysis/feature_viz.py`` -> ``wsae_feature_topk_update``) never does.

Follows /root/reference/src/whisper_sae/analysis/feature_viz.py:
  :94-158   TopKTracker.update      - loop over samples, positions, active features (value > 0);
                                      push while the feature's min-heap has < k entries, else replace
                                      the minimum only if the new value is STRICTLY larger
  :160-172  get_top_examples        - the heap's entries, strongest first
  :182-206  get_feature_stats
Pinned by tests/golden/tracker.pt, generated from the live reference by
oracle/make_golden_tracker.py.

Exact ties: the reference's heap holds ``(value, FeatureActivation)`` tuples, so two equal values
make ``heapq`` compare the (unordered) dataclasses and raise TypeError - ties are outside the
reference's defined behaviour.  This oracle (and the CUDA path) extend the strict ``>`` rule: of
equal values the EARLIER arrival stays, i.e. the latest arrival is the first to be evicted.
"""

from __future__ import annotations

import heapq

import numpy as np


class TrackerOracle:
    def __init__(self, num_features: int, k: int = 20):
        self.num_features = num_features
        self.k = k
        # min-heap per feature of (value, -arrival, sample_idx, position_idx)
        self.heaps: list[list[tuple[float, int, int, int]]] = [[] for _ in range(num_features)]
        self.total_activations = 0
        self.samples_processed = 0
        self._arrival = 0

    def update(self, activations, sample_indices) -> None:
        """``activations``: array-like [batch, features] or [batch, seq, features] (feature_viz.py:107-156)."""
        acts = np.asarray(activations, dtype=np.float32)
        if acts.ndim == 2:
            acts = acts[:, None, :]
        batch, seq, feats = acts.shape
        assert feats == self.num_features
        for b in range(batch):
            sample = int(sample_indices[b])
            for pos in range(seq):
                row = acts[b, pos]
                for f in np.nonzero(row > 0)[0].tolist():
                    v = float(row[f])
                    self.total_activations += 1
                    self._arrival += 1
                    heap = self.heaps[f]
                    entry = (v, -self._arrival, sample, pos)
                    if len(heap) < self.k:
                        heapq.heappush(heap, entry)
                    elif v > heap[0][0]:
                        heapq.heapreplace(heap, entry)
        self.samples_processed += batch

    def update_sparse(self, idx, val, sample_indices) -> None:
        """Same, from a [batch, k] TopK code (signed values; only val > 0 fires, sae/model.py:116)."""
        idx, val = np.asarray(idx), np.asarray(val, dtype=np.float32)
        dense = np.zeros((idx.shape[0], self.num_features), dtype=np.float32)
        for b in range(idx.shape[0]):
            for j in range(idx.shape[1]):
                if idx[b, j] >= 0 and val[b, j] > 0:
                    dense[b, idx[b, j]] = val[b, j]
        self.update(dense, sample_indices)

    def top(self, f: int) -> list[tuple[float, int, int]]:
        """(value, sample_idx, position_idx), strongest first; equal values: earlier arrival first."""
        return [(v, s, p) for v, _, s, p in sorted(self.heaps[f], key=lambda e: (-e[0], -e[1]))]

    def stats(self, f: int) -> dict:
        vals = [e[0] for e in self.top(f)]
        if not vals:
            return {"num_examples": 0, "max_activation": 0.0, "min_activation": 0.0, "mean_activation": 0.0}
        return {"num_examples": len(vals), "max_activation": max(vals), "min_activation": min(vals),
                "mean_activation": sum(vals) / len(vals)}
