"""Top-k activating examples per feature (reference: analysis/feature_viz.py).

Same names and results as the reference's ``FeatureActivation`` / ``TopKTracker`` /
``FeatureReport`` / ``collect_top_activations``; the tracker's state lives in device tensors and a
batch is merged by ``wsae_feature_topk_update`` (four kernels) instead of one ``heapq`` operation
per (sample, position, active feature) in Python (feature_viz.py:107-156).  ``collect_top_activations``
feeds it the ``(idx, val)`` TopK code straight from K1, so the dense ``[B, F]`` ``hidden`` tensor the
reference builds with ``model.encode`` (feature_viz.py:461-462) never exists.

CUDA only: tensors on another device raise (no CPU fallback; the CPU semantics are the reference's).
"""

from __future__ import annotations

import json
from dataclasses import asdict, dataclass, field
from pathlib import Path
from typing import Any

import torch
from torch import Tensor

from .. import ops

FRAME_MS = 10.0   # Whisper encoder frame period used for timestamps (feature_viz.py:131)
MAX_TRACKED = 32  # list slots per feature supported by the merge kernel (one per warp lane)


@dataclass
class FeatureActivation:
    """One firing of one feature (feature_viz.py:22-56)."""

    feature_idx: int
    activation_value: float
    sample_idx: int
    position_idx: int
    timestamp_ms: float | None = None
    transcription: str | None = None
    transcription_context: str | None = None
    audio_path: str | None = None
    metadata: dict[str, Any] = field(default_factory=dict)

    def to_dict(self) -> dict:
        return asdict(self)

    @classmethod
    def from_dict(cls, d: dict) -> "FeatureActivation":
        return cls(**d)


class TopKTracker:
    """The k strongest firings of every feature over a stream of batches (feature_viz.py:59-260).

    Device state: ``top_val [F, k]`` (descending, ``-inf`` = empty), ``top_sample [F, k]`` int64,
    ``top_pos [F, k]`` int32, ``top_count [F]``.  Transcriptions / metadata stay on the host, keyed
    by sample index, and are attached when examples are read back.
    """

    def __init__(self, num_features: int, k: int = 20, device: torch.device | str = "cuda"):
        if not 1 <= k <= MAX_TRACKED:
            raise ValueError(f"k must be in [1, {MAX_TRACKED}] (got {k})")
        self.num_features = num_features
        self.k = k
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("TopKTracker keeps its state on a CUDA device (no CPU fallback)")
        F = num_features
        self.top_val = torch.full((F, k), float("-inf"), dtype=torch.float32, device=self.device)
        self.top_sample = torch.full((F, k), -1, dtype=torch.int64, device=self.device)
        self.top_pos = torch.full((F, k), -1, dtype=torch.int32, device=self.device)
        self.top_count = torch.zeros(F, dtype=torch.int32, device=self.device)
        self._total = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._total_loaded = 0
        self.samples_processed = 0
        self._transcriptions: dict[int, str] = {}
        self._metadata: dict[int, dict] = {}

    # ---- counters -------------------------------------------------------------------------------
    @property
    def total_activations(self) -> int:
        """Number of (sample, position, feature) firings seen (feature_viz.py:128)."""
        return self._total_loaded + int(self._total.item())

    # ---- updates --------------------------------------------------------------------------------
    def _remember(self, sample_indices: list[int], transcriptions, metadata_list) -> None:
        if transcriptions:
            for s, t in zip(sample_indices, transcriptions):
                self._transcriptions[int(s)] = t
        if metadata_list:
            for s, m in zip(sample_indices, metadata_list):
                if m:
                    self._metadata[int(s)] = dict(m)

    @staticmethod
    def _as_index_list(sample_indices) -> list[int]:
        return sample_indices.tolist() if isinstance(sample_indices, Tensor) else list(sample_indices)

    def update(self, activations: Tensor, sample_indices: list[int] | Tensor,
               transcriptions: list[str] | None = None, metadata_list: list[dict] | None = None) -> None:
        """Dense ``[batch, features]`` or ``[batch, seq, features]`` activations (feature_viz.py:94-158)."""
        if not activations.is_cuda:
            raise RuntimeError("TopKTracker.update needs CUDA activations (no CPU fallback)")
        acts = activations.detach().to(torch.float32)
        if acts.ndim == 2:
            acts = acts.unsqueeze(1)
        batch, seq, feats = acts.shape
        assert feats == self.num_features
        ids = self._as_index_list(sample_indices)
        flat = acts.reshape(batch * seq, feats)
        coords = (flat > 0).nonzero()                      # row-major order: row, then feature
        rows = coords[:, 0].to(torch.int32).contiguous()
        feat = coords[:, 1].to(torch.int32).contiguous()
        val = flat[coords[:, 0], coords[:, 1]].contiguous()
        sid = torch.tensor(ids, dtype=torch.int64, device=acts.device).repeat_interleave(seq)
        pos = torch.arange(seq, dtype=torch.int32, device=acts.device).repeat(batch)
        ops.feature_topk_update(feat, val, rows, 1, sid, 0, pos, self.top_val, self.top_sample,
                                self.top_pos, self.top_count, self._total)
        self._remember(ids, transcriptions, metadata_list)
        self.samples_processed += batch

    def update_sparse(self, idx: Tensor, val: Tensor, sample_indices: list[int] | Tensor | int,
                      transcriptions: list[str] | None = None,
                      metadata_list: list[dict] | None = None) -> None:
        """``(idx, val) [batch, k]`` TopK code of one batch (signed pre-activations; only ``val > 0``
        fires, sae/model.py:116).  ``sample_indices``: one id per row, or the id of row 0."""
        batch, k = idx.shape
        if isinstance(sample_indices, int):
            sid, base, ids = None, sample_indices, None
        else:
            ids = self._as_index_list(sample_indices)
            sid, base = torch.tensor(ids, dtype=torch.int64, device=idx.device), 0
        ops.feature_topk_update(idx.contiguous().view(-1), val.contiguous().view(-1), None, k, sid, base,
                                None, self.top_val, self.top_sample, self.top_pos, self.top_count,
                                self._total)
        if ids is not None:
            self._remember(ids, transcriptions, metadata_list)
        self.samples_processed += batch

    # ---- read-back ------------------------------------------------------------------------------
    def _example(self, feature_idx: int, value: float, sample: int, pos: int) -> FeatureActivation:
        return FeatureActivation(
            feature_idx=feature_idx, activation_value=value, sample_idx=sample, position_idx=pos,
            timestamp_ms=pos * FRAME_MS, transcription=self._transcriptions.get(sample),
            metadata=dict(self._metadata.get(sample, {})))

    def get_top_examples(self, feature_idx: int) -> list[FeatureActivation]:
        """Strongest first (feature_viz.py:160-172)."""
        n = int(self.top_count[feature_idx])
        vals = self.top_val[feature_idx, :n].tolist()
        samples = self.top_sample[feature_idx, :n].tolist()
        poss = self.top_pos[feature_idx, :n].tolist()
        return [self._example(feature_idx, v, s, p) for v, s, p in zip(vals, samples, poss)]

    def get_all_top_examples(self) -> dict[int, list[FeatureActivation]]:
        counts = self.top_count.tolist()
        vals, samples, poss = self.top_val.tolist(), self.top_sample.tolist(), self.top_pos.tolist()
        return {f: [self._example(f, vals[f][j], samples[f][j], poss[f][j]) for j in range(counts[f])]
                for f in range(self.num_features)}

    def get_feature_stats(self) -> dict[int, dict]:
        """num_examples / max / min / mean of each feature's kept values (feature_viz.py:182-206)."""
        counts = self.top_count.to(torch.float64)
        filled = torch.arange(self.k, device=self.device)[None, :] < self.top_count[:, None]
        vals = torch.where(filled, self.top_val, torch.zeros_like(self.top_val)).to(torch.float64)
        means = (vals.sum(1) / counts.clamp(min=1)).tolist()
        maxs = vals[:, 0].tolist()                                   # lists are kept descending
        last = (self.top_count.long() - 1).clamp(min=0)
        mins = vals.gather(1, last[:, None])[:, 0].tolist()
        out = {}
        for f, n in enumerate(self.top_count.tolist()):
            out[f] = ({"num_examples": n, "max_activation": maxs[f], "min_activation": mins[f],
                       "mean_activation": means[f]} if n else
                      {"num_examples": 0, "max_activation": 0.0, "min_activation": 0.0,
                       "mean_activation": 0.0})
        return out

    # ---- persistence (same JSON schema as feature_viz.py:208-260) ---------------------------------
    def save(self, path: Path | str) -> None:
        examples = self.get_all_top_examples()
        data = {"num_features": self.num_features, "k": self.k,
                "total_activations": self.total_activations,
                "samples_processed": self.samples_processed,
                "features": {str(f): [e.to_dict() for e in ex] for f, ex in examples.items() if ex}}
        Path(path).write_text(json.dumps(data, indent=2))

    @classmethod
    def load(cls, path: Path | str, device: torch.device | str = "cuda") -> "TopKTracker":
        data = json.loads(Path(path).read_text())
        tr = cls(num_features=data["num_features"], k=data["k"], device=device)
        tr._total_loaded = data["total_activations"]
        tr.samples_processed = data["samples_processed"]
        val, smp, pos, cnt = tr.top_val.cpu(), tr.top_sample.cpu(), tr.top_pos.cpu(), tr.top_count.cpu()
        for f_str, examples in data["features"].items():
            f = int(f_str)
            acts = sorted((FeatureActivation.from_dict(e) for e in examples),
                          key=lambda e: e.activation_value, reverse=True)[: tr.k]
            for j, e in enumerate(acts):
                val[f, j], smp[f, j], pos[f, j] = e.activation_value, e.sample_idx, e.position_idx
                if e.transcription is not None:
                    tr._transcriptions[e.sample_idx] = e.transcription
                if e.metadata:
                    tr._metadata[e.sample_idx] = dict(e.metadata)
            cnt[f] = len(acts)
        tr.top_val.copy_(val), tr.top_sample.copy_(smp), tr.top_pos.copy_(pos), tr.top_count.copy_(cnt)
        return tr


@dataclass
class FeatureInterpretation:
    """What a feature is believed to represent (feature_viz.py:262-281)."""

    feature_idx: int
    category: str
    description: str
    confidence: float
    evidence: list[str] = field(default_factory=list)
    automated_labels: dict[str, Any] = field(default_factory=dict)

    def to_dict(self) -> dict:
        return asdict(self)


class FeatureReport:
    """JSON reports over a tracker (feature_viz.py:284-422): summary.json, features/feature_NNNNN.json,
    tracker_state.json."""

    def __init__(self, tracker: TopKTracker, output_dir: Path | str):
        self.tracker = tracker
        self.output_dir = Path(output_dir)
        self.output_dir.mkdir(parents=True, exist_ok=True)
        self.interpretations: dict[int, FeatureInterpretation] = {}

    def generate_feature_report(self, feature_idx: int, include_audio_paths: bool = True) -> dict:
        tops = []
        for ex in self.tracker.get_top_examples(feature_idx):
            row = {key: getattr(ex, key) for key in
                   ("activation_value", "sample_idx", "position_idx", "timestamp_ms", "transcription")}
            if include_audio_paths and ex.audio_path:
                row["audio_path"] = ex.audio_path
            tops.append(row)
        report = {"feature_idx": feature_idx, "stats": self.tracker.get_feature_stats()[feature_idx],
                  "top_examples": tops}
        if feature_idx in self.interpretations:
            report["interpretation"] = self.interpretations[feature_idx].to_dict()
        return report

    def generate_summary_report(self, top_n: int = 100) -> dict:
        stats = self.tracker.get_feature_stats()
        ranked = sorted(stats.items(), key=lambda kv: kv[1]["max_activation"], reverse=True)[:top_n]
        return {"num_features": self.tracker.num_features,
                "samples_processed": self.tracker.samples_processed,
                "total_activations": self.tracker.total_activations,
                "top_features": [{"feature_idx": f, **st} for f, st in ranked]}

    def save_reports(self, top_n: int = 100) -> None:
        summary = self.generate_summary_report(top_n=top_n)
        (self.output_dir / "summary.json").write_text(json.dumps(summary, indent=2))
        features_dir = self.output_dir / "features"
        features_dir.mkdir(exist_ok=True)
        for entry in summary["top_features"]:
            f = entry["feature_idx"]
            (features_dir / f"feature_{f:05d}.json").write_text(
                json.dumps(self.generate_feature_report(f), indent=2))
        self.tracker.save(self.output_dir / "tracker_state.json")

    def add_interpretation(self, feature_idx: int, category: str, description: str,
                           confidence: float = 0.5, evidence: list[str] | None = None) -> None:
        self.interpretations[feature_idx] = FeatureInterpretation(
            feature_idx=feature_idx, category=category, description=description,
            confidence=confidence, evidence=evidence or [])


def collect_top_activations(model: torch.nn.Module, dataloader, num_features: int, k: int = 20,
                            device: str = "cuda") -> TopKTracker:
    """Run ``model`` over ``dataloader`` and track every feature's top-k firings
    (feature_viz.py:425-484).  Models with a fused sparse encoder (``TopKSAE``) hand their
    ``(idx, val)`` code to the tracker; anything else goes through its dense ``encode``/``forward``."""
    tracker = TopKTracker(num_features=num_features, k=k, device=device)
    model.eval()
    sample_idx = 0
    with torch.no_grad():
        for batch in dataloader:
            metadata = None
            if isinstance(batch, (tuple, list)):
                activations = batch[0]
                metadata = batch[1] if len(batch) > 1 else None
            else:
                activations = batch
            activations = activations.to(device)
            transcriptions = metadata.get("transcriptions") if isinstance(metadata, dict) else None
            n = activations.shape[0]
            ids = list(range(sample_idx, sample_idx + n))
            sparse = getattr(model, "_sparse_encode", None)
            if sparse is not None and activations.ndim == 2:
                idx, val = sparse(activations)
                tracker.update_sparse(idx, val, ids, transcriptions=transcriptions)
            else:
                if hasattr(model, "encode"):
                    hidden = model.encode(activations)
                else:
                    output = model(activations)
                    hidden = output.hidden if hasattr(output, "hidden") else output[1]
                tracker.update(hidden, ids, transcriptions=transcriptions)
            sample_idx += n
    return tracker
