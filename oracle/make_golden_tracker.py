"""Golden vectors for the per-feature top-k tracker from the LIVE reference (build container only:
/root/reference does not exist on the GPU box).  Writes tests/golden/tracker.pt.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_tracker.py
"""

from __future__ import annotations

import sys
from pathlib import Path

import torch

REF_SRC = Path("/root/reference/src")
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF_SRC))
sys.dont_write_bytecode = True

from whisper_sae.analysis.feature_viz import TopKTracker  # noqa: E402  (reference)

GOLDEN = ROOT / "tests" / "golden"


def sparse_batch(gen: torch.Generator, rows: int, F: int, k_active: int, seq: int | None) -> torch.Tensor:
    """Dense activations with k_active distinct positive entries per (row, position) plus a few
    non-positive ones that must not count; continuous values, so no exact ties."""
    shape = (rows, F) if seq is None else (rows, seq, F)
    flat_rows = rows * (seq or 1)
    acts = torch.zeros(flat_rows, F)
    for r in range(flat_rows):
        cols = torch.randperm(F, generator=gen)[: k_active + 2]
        acts[r, cols[:k_active]] = torch.rand(k_active, generator=gen) * 3 + 1e-3
        acts[r, cols[k_active:]] = -torch.rand(2, generator=gen)        # negative: never fires
    return acts.reshape(shape)


def case(name: str, F: int, k: int, batches: list[tuple[int, int | None]], k_active: int, seed: int) -> dict:
    gen = torch.Generator().manual_seed(seed)
    ref = TopKTracker(num_features=F, k=k)
    inputs, ids, next_id = [], [], 0
    for rows, seq in batches:
        acts = sparse_batch(gen, rows, F, k_active, seq)
        sample_ids = list(range(next_id, next_id + rows))
        next_id += rows
        ref.update(acts, sample_ids)
        inputs.append(acts)
        ids.append(sample_ids)
    expected = {f: [(e.activation_value, e.sample_idx, e.position_idx, e.timestamp_ms)
                    for e in ref.get_top_examples(f)] for f in range(F)}
    return {"recipe": dict(name=name, F=F, k=k, batches=batches, k_active=k_active, seed=seed),
            "inputs": inputs, "sample_ids": ids, "expected": expected, "stats": ref.get_feature_stats(),
            "total_activations": ref.total_activations, "samples_processed": ref.samples_processed}


def main() -> None:
    fixtures = {
        # several updates, lists overflow (k = 5 < firings per feature), 2-D inputs
        "flat_64": case("flat_64", F=64, k=5, batches=[(48, None), (48, None), (32, None)], k_active=8, seed=1),
        # 3-D inputs: position index and timestamp
        "seq_32": case("seq_32", F=32, k=4, batches=[(6, 7), (5, 7)], k_active=4, seed=2),
        # k larger than the number of firings of most features
        "sparse_128": case("sparse_128", F=128, k=20, batches=[(40, None), (24, None)], k_active=6, seed=3),
    }
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.save({"torch": torch.__version__, "cases": fixtures}, GOLDEN / "tracker.pt")
    for n, fx in fixtures.items():
        print(n, "total_activations", fx["total_activations"], "samples", fx["samples_processed"])


if __name__ == "__main__":
    main()
