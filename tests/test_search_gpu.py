"""GPU parity of the search half through the C ABI (drag_topk & friends) against the
reference-generated goldens and the oracle."""

import numpy as np
import pytest

from oracle import search as osearch
from tests.helpers import load_json, synth_cases, unit_docs
from tests.synth import synth_matrix, synth_queries

pytestmark = pytest.mark.gpu


def _index(docs, metric, limit, **kw):
    from dial_rag_b200.records import RetrievalType
    from dial_rag_b200.retrievers.embeddings_index import DocIndex, EmbeddingsIndex, Metric

    return EmbeddingsIndex(RetrievalType.TEXT, [DocIndex(c, e) for c, e in docs], metric=Metric(metric), limit=limit, **kw)


def _pairs(docs):
    return [[d.metadata["doc_id"], d.metadata["chunk_id"]] for d in docs]


def test_native_library_is_loaded():
    from dial_rag_b200 import _native

    info = _native.device_info(0)
    assert info["compute_capability"] == 100, info
    assert info["sm_count"] >= 100


def test_reference_unit_fixtures():
    from dial_rag_b200.records import RetrievalType, to_metadata_doc

    for case in load_json("search_unit.json"):
        idx = _index(unit_docs(case["order"]), case["metric"], case["limit"])
        got = idx.find(np.array(case["query"]))
        assert _pairs(got) == case["expected"], case
        assert got == [to_metadata_doc(d, c, RetrievalType.TEXT) for d, c in case["expected"]]


def test_row_sqnorm_matches_numpy_bitwise():
    from dial_rag_b200.device_index import DeviceMatrix

    for dim in (3, 7, 16, 100, 129, 384, 1024, 1408):
        m = synth_matrix(seed=dim, rows=257, dim=dim, normalise=dim % 2 == 0)
        dm = DeviceMatrix(m)
        assert np.array_equal(dm.row_sq.cpu().numpy(), np.sum(m**2, axis=1)), dim


def test_distances_match_reference_values():
    from dial_rag_b200.retrievers.embeddings_metrics import ENUM_TO_METRIC, Metric

    kat = load_json("search_metrics_kat.json")
    docs = np.array(kat["docs"], dtype=np.float32)
    q = np.array(kat["query"], dtype=np.float64)
    for metric, hexes in kat["distances"].items():
        want = np.array([float.fromhex(h) for h in hexes])
        got = ENUM_TO_METRIC[Metric(metric)](q, docs)
        assert got.dtype == np.float64
        assert np.array_equal(np.isnan(got), np.isnan(want)), metric
        # cosine: the reference normalises the float32 rows IN float32 (torch vectorised norm, whose
        # summation order depends on the host SIMD width) before promoting to float64, so its values
        # carry ~1e-7 of float32 noise; ours are exact float64.  Other metrics agree to float64 rounding.
        tol = dict(rtol=0, atol=2e-7) if metric == "cosine_sim" else dict(rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(got, want, equal_nan=True, **tol)
    got = ENUM_TO_METRIC[Metric.COSINE_SIM](np.zeros(16), docs)
    assert np.all(got == 0.0)


def test_reference_metric_known_answers():
    # reference tests/test_embeddings_metrics.py, default assert_allclose tolerance
    from dial_rag_b200.retrievers.embeddings_metrics import ENUM_TO_METRIC, Metric

    a = np.array
    f = {m: ENUM_TO_METRIC[m] for m in Metric}
    e4 = a([[1.0, 0, 0, 0], [0, 1.0, 0, 0]])
    np.testing.assert_allclose(f[Metric.COSINE_SIM](a([1.0, 0, 0, 0]), e4), [-1.0, 0.0])
    np.testing.assert_allclose(f[Metric.COSINE_SIM](a([-1.0, 0, 0, 0]), e4), [1.0, 0.0])
    np.testing.assert_allclose(f[Metric.COSINE_SIM](a([2.0, 0, 0, 0]), e4), [-1.0, 0.0])
    z3 = a([[1.0, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 0]])
    np.testing.assert_allclose(f[Metric.COSINE_SIM](a([0.0, 0, 0, 0]), z3), [0.0, 0.0, 0.0])
    np.testing.assert_allclose(f[Metric.COSINE_SIM](a([1.0, 0, 0, 0]), 2 * z3), [-1.0, 0.0, 0.0])
    np.testing.assert_allclose(f[Metric.INNER_PRODUCT](a([2, 0, 0, 0]), e4), [-2.0, 0.0])
    np.testing.assert_allclose(f[Metric.INNER_PRODUCT](a([1, 0, 0, 0]), 2 * z3), [-2.0, 0.0, 0.0])
    np.testing.assert_allclose(f[Metric.EUCLIDEAN_DIST](a([1, 0, 0, 0]), e4), [0.0, np.sqrt(2)])
    np.testing.assert_allclose(f[Metric.EUCLIDEAN_DIST](a([2, 0, 0, 0]), e4), [1.0, np.sqrt(5)])
    np.testing.assert_allclose(
        f[Metric.EUCLIDEAN_DIST](a([1, 0, 0, 0]), a([[2, 0, 0, 0], [3, 3, 3, 0], [0, 0, 0, 0]])), [1.0, np.sqrt(22), 1.0])
    np.testing.assert_allclose(f[Metric.SQEUCLIDEAN_DIST](a([-1, 0, 0, 0]), e4), [4.0, 2.0])
    np.testing.assert_allclose(f[Metric.SQEUCLIDEAN_DIST](a([0, 0, 0, 0]), a([[1, 1, 1, 1], [2, 2, 2, 2]])), [4.0, 16.0])
    q = a([1, 2, 3, 4])
    docs = a([[1, 0, 0, 0], [0, 1, 0, 0], [2, 0, 0, 0], [3, 3, 3, 0], [0, 0, 0, 0]])
    np.testing.assert_allclose(f[Metric.EUCLIDEAN_DIST](q, docs) ** 2, f[Metric.SQEUCLIDEAN_DIST](q, docs))
    n = lambda x: x / np.maximum(np.linalg.norm(x, axis=-1, keepdims=True), 1e-30)  # noqa: E731
    np.testing.assert_allclose(f[Metric.COSINE_SIM](n(q.astype(float)), n(docs.astype(float))),
                               f[Metric.INNER_PRODUCT](n(q.astype(float)), n(docs.astype(float))), atol=1e-15)


def test_synth_cases_bit_exact_ids_vs_reference():
    """(doc_id, chunk_id) lists identical to what the real reference returned."""
    checked = 0
    for entry, data in synth_cases():
        for metric in osearch.ALL_METRICS:
            for limit in (1, 7, 100):
                idx = _index(data["docs"], metric, limit)
                want = [r for r in entry["results"] if r["metric"] == metric and r["limit"] == limit]
                got_batch = idx.find_batch(data["queries"])
                for r in want:
                    one = idx.find(data["queries"][r["query"]])
                    assert _pairs(one) == r["expected"], (entry["name"], metric, limit, r["query"], "single")
                    assert _pairs(got_batch[r["query"]]) == r["expected"], (entry["name"], metric, limit, r["query"], "batch")
                    checked += 1
        big = entry["in_doc"]["doc"]
        for row in entry["in_doc"]["rows"]:
            idx = _index([], row["metric"], 20)
            from dial_rag_b200.retrievers.embeddings_index import DocIndex

            ids, dist = idx.find_in_doc(data["queries"][0], DocIndex(*data["docs"][big]))
            assert ids.tolist() == row["chunk_ids"], (entry["name"], row["metric"])
            want_d = np.array([float.fromhex(h) for h in row["distances"]])
            tol = dict(rtol=0, atol=2e-7) if row["metric"] == "cosine_sim" else dict(rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(dist, want_d, equal_nan=True, **tol)
    assert checked >= 300


@pytest.mark.parametrize("metric", osearch.ALL_METRICS)
@pytest.mark.parametrize("k", [1, 20, 100, 1000])
def test_topk_rows_vs_oracle_200k(metric, k):
    """Seeded 200k x 384 matrix with planted duplicates, ids bit-exact vs the oracle."""
    from dial_rag_b200.device_index import DeviceMatrix

    m = synth_matrix(seed=2, rows=200_003, dim=384)
    rng = np.random.Generator(np.random.PCG64(9))
    dst = rng.integers(0, len(m), size=3000)
    m[dst] = m[rng.integers(0, len(m), size=3000)]
    q = synth_queries(seed=3, n=5, dim=384)
    q[0] = m[dst[0]].astype(np.float64)  # exact hit with duplicates
    dm = DeviceMatrix(m)
    dist, rows, count = dm.topk(q, k, metric)
    assert count.tolist() == [k] * len(q)
    for i in range(len(q)):
        want_rows, want_d = osearch.topk_rows(metric, k, q[i], m)
        assert np.array_equal(rows[i], want_rows), (metric, k, i)
        tol = dict(rtol=0, atol=2e-7) if metric == "cosine_sim" else dict(rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(dist[i], want_d, equal_nan=True, **tol)


def test_limit_larger_than_rows_and_tiny_dims():
    from dial_rag_b200.device_index import DeviceMatrix

    m = synth_matrix(seed=5, rows=37, dim=5, normalise=False)
    q = synth_queries(seed=6, n=3, dim=5, normalise=False)
    dm = DeviceMatrix(m)
    for metric in osearch.ALL_METRICS:
        dist, rows, count = dm.topk(q, 100, metric)
        assert rows.shape == (3, 37) and count.tolist() == [37] * 3
        for i in range(3):
            want_rows, _ = osearch.topk_rows(metric, 100, q[i], m)
            assert np.array_equal(rows[i], want_rows)


def test_all_identical_rows_tie_break_lowest_id():
    from dial_rag_b200.device_index import DeviceMatrix

    m = np.tile(synth_matrix(seed=8, rows=1, dim=384), (50_000, 1))
    q = synth_queries(seed=9, n=2, dim=384)
    dm = DeviceMatrix(m)
    for metric in osearch.ALL_METRICS:
        _, rows, _ = dm.topk(q, 100, metric)
        assert np.array_equal(rows[0], np.arange(100)) and np.array_equal(rows[1], np.arange(100)), metric
    z = np.zeros((10_000, 384), dtype=np.float32)
    _, rows, _ = DeviceMatrix(z).topk(np.zeros((1, 384)), 7, "cosine_sim")
    assert rows[0].tolist() == list(range(7))


def test_bf16_storage_recall_and_exactness():
    """bf16 path: exact top-k of the bf16-rounded values; recall@100 vs fp32 truth >= 0.999 is
    BASELINE's bar on the *stored* values, score tolerance 2^-8."""
    import torch

    from dial_rag_b200.device_index import DeviceMatrix

    m = synth_matrix(seed=4, rows=100_000, dim=384)
    q = synth_queries(seed=14, n=4, dim=384)
    rounded = torch.from_numpy(m).to(torch.bfloat16).to(torch.float32).numpy()
    dm = DeviceMatrix(m, storage="bf16")
    dist, rows, _ = dm.topk(q, 100, "inner_product")
    for i in range(len(q)):
        want_rows, want_d = osearch.topk_rows("inner_product", 100, q[i], rounded)
        assert len(set(rows[i]) & set(want_rows)) >= 100 * 0.999
        assert np.array_equal(rows[i], want_rows)
        np.testing.assert_allclose(dist[i], want_d, rtol=1e-11, atol=1e-12)
        full = -(m.astype(np.float64) @ q[i])
        assert np.max(np.abs(full[rows[i]] - dist[i])) <= 2.0**-8
    # one query at a time goes through the shared-memory ring scan on the bf16 rows (the four above took the
    # tensor-core candidate pass + float64 re-rank): same float64 arithmetic, bit-identical answers
    for i in range(2):
        d1, r1, _ = dm.topk(q[i:i + 1], 100, "inner_product")
        assert np.array_equal(r1[0], rows[i]) and np.array_equal(d1[0], dist[i])


def test_sharded_merge_equals_single_scan():
    """Row-sharded index: per-shard top-k + drag_topk_merge == one scan."""
    import torch

    from dial_rag_b200 import _native
    from dial_rag_b200.device_index import DeviceMatrix, merge_topk_device

    m = synth_matrix(seed=12, rows=60_000, dim=384)
    m[40_000:40_050] = m[100:150]  # ties across shards
    q = synth_queries(seed=13, n=6, dim=384)
    q[1] = m[120].astype(np.float64)
    k = 50
    bounds = [0, 7_001, 7_001, 33_000, 60_000]  # includes an empty shard
    whole = DeviceMatrix(m)
    for metric in ("inner_product", "sqeuclidean_dist"):
        d_all, r_all, _ = whole.topk(q, k, metric)
        dq = torch.from_numpy(q).cuda()
        parts = []
        for s in range(len(bounds) - 1):
            shard = DeviceMatrix(m[bounds[s]:bounds[s + 1]], row_id_base=bounds[s])
            if shard.n_rows == 0:
                parts.append((torch.full((len(q), k), float("nan"), dtype=torch.float64, device="cuda"),
                              torch.full((len(q), k), -1, dtype=torch.int64, device="cuda"),
                              torch.zeros(len(q), dtype=torch.int32, device="cuda")))
            else:
                parts.append(shard.topk_device(dq, k, metric))
        dist = torch.stack([p[0] for p in parts])
        rows = torch.stack([p[1] for p in parts])
        cnt = torch.stack([p[2] for p in parts])
        md, mr, mc = merge_topk_device(_native.load(), 0, dist, rows, cnt, k)
        assert np.array_equal(mr.cpu().numpy(), r_all), metric
        assert np.array_equal(md.cpu().numpy(), d_all), metric
        assert mc.cpu().tolist() == [k] * len(q)


def test_errors_are_python_exceptions():
    from dial_rag_b200._native import DragError
    from dial_rag_b200.device_index import DeviceMatrix

    dm = DeviceMatrix(synth_matrix(seed=1, rows=5000, dim=384))
    with pytest.raises(ValueError):
        dm.topk(np.zeros((1, 100)), 5, "inner_product")
    # k beyond the kernels' list capacity is not an error (the reference takes any limit): it goes through the
    # distances + stable sort path (tests/test_scale_and_threads_gpu.py checks its ids)
    dist, rows, count = dm.topk(np.zeros((1, 384)), 5000, "inner_product")
    assert count.tolist() == [5000] and np.array_equal(rows[0][:5000], np.arange(5000))   # all scores equal: row order
    dist, rows, count = dm.topk(np.zeros((1, 384)), 0, "inner_product")   # limit 0: empty, as argsort()[:0]
    assert rows.shape == (1, 0) and count.tolist() == [0]
    import torch

    with pytest.raises(DragError):   # the device-tensor entry point validates k itself
        dm.topk_device(torch.zeros((1, 384), dtype=torch.float64, device="cuda"), 0, "inner_product")
    with pytest.raises(ValueError):
        dm.topk(np.zeros((1, 384)), 5, "manhattan")


@pytest.mark.parametrize("dim", [1024, 1408])
@pytest.mark.parametrize("metric", ["cosine_sim", "sqeuclidean_dist"])
def test_wide_rows_many_queries(dim, metric):
    """D = 1024 / 1408 (the description and multimodal retrievers' embeddings) with more queries than the
    batched path's threshold: beyond its D <= 512 they must go through the float64 scan, ids bit-exact vs the oracle."""
    from dial_rag_b200.device_index import DeviceMatrix

    m = synth_matrix(seed=dim, rows=70_001, dim=dim, normalise=False)
    q = synth_queries(seed=dim + 1, n=9, dim=dim)
    dm = DeviceMatrix(m)
    dist, rows, count = dm.topk(q, 20, metric)
    assert count.tolist() == [20] * len(q)
    for i in range(len(q)):
        want_rows, want_d = osearch.topk_rows(metric, 20, q[i], m)
        assert np.array_equal(rows[i], want_rows), (dim, metric, i)
        tol = dict(rtol=0, atol=2e-7) if metric == "cosine_sim" else dict(rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(dist[i], want_d, equal_nan=True, **tol)
