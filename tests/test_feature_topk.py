"""Per-feature top-k tracker (analysis/feature_viz.py of the reference).

CPU: the oracle restatement against golden vectors produced by the live reference's TopKTracker
(oracle/make_golden_tracker.py -> tests/golden/tracker.pt).
GPU: whisper_sae_b200.analysis.TopKTracker (wsae_feature_topk_update through the C ABI) against the
same fixtures and against the oracle on cases the fixtures do not hold (exact ties, the sparse
[B, k] entry point, full-size batches).  Values are compared bit-exactly: the kernels only move
them.  The GPU tests read like the reference's tests/test_analysis.py::TestTopKTracker.
"""

import tempfile
from pathlib import Path

import pytest
import torch

from oracle.feature_topk_oracle import TrackerOracle

GOLDEN = Path(__file__).parent / "golden" / "tracker.pt"
CASES = ["flat_64", "seq_32", "sparse_128"]


def _fixture(name):
    return torch.load(GOLDEN, weights_only=False)["cases"][name]


# ---------------------------------------------------------------- CPU: oracle pinned to the reference
@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    fx = _fixture(name)
    r = fx["recipe"]
    o = TrackerOracle(r["F"], r["k"])
    for acts, ids in zip(fx["inputs"], fx["sample_ids"]):
        o.update(acts.numpy(), ids)
    assert o.total_activations == fx["total_activations"]
    assert o.samples_processed == fx["samples_processed"]
    for f in range(r["F"]):
        want = [(v, s, p) for v, s, p, _ in fx["expected"][f]]
        assert o.top(f) == want, f"feature {f}"
        assert o.stats(f)["num_examples"] == fx["stats"][f]["num_examples"]
        assert o.stats(f)["mean_activation"] == pytest.approx(fx["stats"][f]["mean_activation"], rel=1e-12)


def test_oracle_tie_rule_keeps_earlier_arrival():
    o = TrackerOracle(4, k=2)
    acts = torch.zeros(4, 4)
    acts[:, 1] = torch.tensor([0.5, 0.5, 0.5, 0.7])
    o.update(acts.numpy(), [10, 11, 12, 13])
    assert o.top(1) == [(pytest.approx(0.7), 13, 0), (0.5, 10, 0)]


# ---------------------------------------------------------------- GPU: the CUDA tracker
def _tracker(F, k):
    from whisper_sae_b200.analysis import TopKTracker
    return TopKTracker(num_features=F, k=k, device="cuda")


def _assert_same(tracker, oracle: TrackerOracle):
    assert tracker.total_activations == oracle.total_activations
    assert tracker.samples_processed == oracle.samples_processed
    everything = tracker.get_all_top_examples()
    for f in range(oracle.num_features):
        got = [(e.activation_value, e.sample_idx, e.position_idx) for e in everything[f]]
        assert got == oracle.top(f), f"feature {f}"


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_tracker_matches_reference_golden(name):
    fx = _fixture(name)
    r = fx["recipe"]
    t = _tracker(r["F"], r["k"])
    for acts, ids in zip(fx["inputs"], fx["sample_ids"]):
        t.update(acts.cuda(), ids)
    assert t.total_activations == fx["total_activations"]
    assert t.samples_processed == fx["samples_processed"]
    stats = t.get_feature_stats()
    for f in range(r["F"]):
        got = [(e.activation_value, e.sample_idx, e.position_idx, e.timestamp_ms) for e in t.get_top_examples(f)]
        assert got == fx["expected"][f], f"feature {f}"
        assert stats[f]["num_examples"] == fx["stats"][f]["num_examples"]
        for key in ("max_activation", "min_activation", "mean_activation"):
            assert stats[f][key] == pytest.approx(fx["stats"][f][key], rel=1e-6)


@pytest.mark.gpu
def test_update_single_sample_and_limit():
    """tests/test_analysis.py:94-157 of the reference, on the device tracker."""
    t = _tracker(64, 5)
    a = torch.zeros(1, 64, device="cuda")
    a[0, 10], a[0, 20] = 0.5, 0.8
    t.update(a, sample_indices=[0])
    assert t.samples_processed == 1 and t.total_activations == 2
    assert [e.activation_value for e in t.get_top_examples(10)] == [0.5]
    assert t.get_top_examples(20)[0].activation_value == pytest.approx(0.8)
    t = _tracker(64, 3)
    for i in range(5):
        a = torch.zeros(1, 64, device="cuda")
        a[0, 0] = i / 10
        t.update(a, sample_indices=[i])
    ex = t.get_top_examples(0)
    assert [e.activation_value for e in ex] == pytest.approx([0.4, 0.3, 0.2])
    assert [e.sample_idx for e in ex] == [4, 3, 2]


@pytest.mark.gpu
def test_sequence_positions_transcriptions_and_roundtrip():
    t = _tracker(64, 10)
    a = torch.zeros(2, 100, 64, device="cuda")
    a[0, 50, 0], a[1, 2, 0], a[1, 4, 10] = 1.0, 0.25, 0.3
    t.update(a, sample_indices=torch.tensor([7, 9]), transcriptions=["hello world", "foo bar"],
             metadata_list=[{"speaker": 1}, {}])
    ex = t.get_top_examples(0)
    assert (ex[0].position_idx, ex[0].timestamp_ms, ex[0].sample_idx) == (50, 500.0, 7)
    assert ex[0].transcription == "hello world" and ex[0].metadata == {"speaker": 1}
    assert ex[1].transcription == "foo bar" and ex[1].position_idx == 2
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "tracker.json"
        t.save(path)
        from whisper_sae_b200.analysis import FeatureReport, TopKTracker
        loaded = TopKTracker.load(path)
        assert (loaded.num_features, loaded.k, loaded.samples_processed) == (64, 10, 2)
        assert loaded.total_activations == 3
        assert [e.to_dict() for e in loaded.get_top_examples(0)] == [e.to_dict() for e in ex]
        # a loaded tracker keeps merging
        b = torch.zeros(1, 64, device="cuda")
        b[0, 0] = 0.5
        loaded.update(b, [11])
        assert [e.sample_idx for e in loaded.get_top_examples(0)] == [7, 11, 9]
        rep = FeatureReport(loaded, Path(tmp) / "reports")
        rep.add_interpretation(0, "phoneme", "test")
        rep.save_reports(top_n=3)
        assert (Path(tmp) / "reports" / "summary.json").exists()
        assert (Path(tmp) / "reports" / "features" / "feature_00000.json").exists()


@pytest.mark.gpu
def test_exact_ties_keep_the_earlier_arrival():
    """Equal values: strict '>' replacement (feature_viz.py:152) - the earlier arrival stays, across
    and within batches."""
    F, k = 8, 3
    t, o = _tracker(F, k), TrackerOracle(F, k)
    batches = [torch.tensor([0.5, 0.5, 0.25, 0.5]), torch.tensor([0.5, 0.75, 0.25, 0.5]),
               torch.tensor([0.75, 0.75, 0.75, 0.75])]
    sid = 0
    for vals in batches:
        a = torch.zeros(len(vals), F)
        a[:, 3] = vals
        a[:, 5] = vals.flip(0)
        ids = list(range(sid, sid + len(vals)))
        sid += len(vals)
        t.update(a.cuda(), ids)
        o.update(a.numpy(), ids)
        _assert_same(t, o)
    assert [e.sample_idx for e in t.get_top_examples(3)] == [5, 8, 9]


@pytest.mark.gpu
@pytest.mark.parametrize("B,F,k_code,K", [(300, 96, 8, 20), (1024, 512, 32, 20), (77, 40, 4, 32), (5, 16, 3, 1)])
def test_sparse_code_entry_point_matches_oracle(B, F, k_code, K):
    """[B, k] TopK codes (signed values, some <= 0, some padded idx = -1) over several batches."""
    g = torch.Generator().manual_seed(B + F)
    t, o = _tracker(F, K), TrackerOracle(F, K)
    sid = 1000
    for step in range(3):
        idx = torch.stack([torch.randperm(F, generator=g)[:k_code] for _ in range(B)]).to(torch.int32)
        val = torch.randn(B, k_code, generator=g)              # about half are <= 0: they never fire
        if step == 1:
            idx[::7, -1] = -1                                  # padded slot of a short row
            val[::7, -1] = float("-inf")
        ids = list(range(sid, sid + B))
        sid += B
        t.update_sparse(idx.cuda(), val.cuda(), ids if step != 2 else ids[0])
        o.update_sparse(idx.numpy(), val.numpy(), ids)
        _assert_same(t, o)


@pytest.mark.gpu
def test_collect_top_activations_uses_the_sparse_code():
    from whisper_sae_b200.analysis import collect_top_activations
    from whisper_sae_b200.sae import TopKSAE
    torch.manual_seed(0)
    sae = TopKSAE(64, 256, k=8).cuda()
    x = torch.randn(96, 64, generator=torch.Generator().manual_seed(1))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x), batch_size=32)
    t = collect_top_activations(sae, loader, num_features=256, k=5, device="cuda")
    o = TrackerOracle(256, 5)
    sae.eval()
    with torch.no_grad():
        o.update(sae.encode(x.cuda()).cpu().numpy(), list(range(96)))   # dense hidden, as the reference does
    _assert_same(t, o)


@pytest.mark.gpu
def test_full_size_batch_properties():
    """B = 75776 rows x k = 32 over F = 3072 (the bench shape): per feature the kept values are the
    K largest positive values of that feature's column, in descending order, with the right rows."""
    B, F, k_code, K = 75776, 3072, 32, 20
    g = torch.Generator(device="cuda").manual_seed(5)
    idx = torch.rand(B, F, device="cuda", generator=g).topk(k_code, dim=1).indices.to(torch.int32)
    val = torch.randn(B, k_code, device="cuda", generator=g)
    t = _tracker(F, K)
    t.update_sparse(idx[: B // 2], val[: B // 2], 0)
    t.update_sparse(idx[B // 2:], val[B // 2:], B // 2)
    dense = torch.full((B, F), float("-inf"), device="cuda")
    dense.scatter_(1, idx.long(), torch.where(val > 0, val, torch.full_like(val, float("-inf"))))
    want_val, want_row = dense.t().topk(K, dim=1)
    assert t.total_activations == int((val > 0).sum())
    assert torch.equal(t.top_count.long(), (dense > float("-inf")).sum(0).clamp(max=K))
    assert torch.equal(t.top_val, want_val)
    filled = want_val > float("-inf")
    assert torch.equal(t.top_sample[filled], want_row[filled])


@pytest.mark.gpu
def test_cpu_tensors_are_rejected():
    t = _tracker(8, 2)
    with pytest.raises(RuntimeError):
        t.update(torch.zeros(1, 8), [0])
