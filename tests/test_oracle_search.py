"""Pins oracle/search.py to outputs of the REAL reference (tests/golden/search_*.json,
made by oracle/make_golden_search.py) and to the known answers of the reference's
tests/test_embeddings_metrics.py and tests/test_embeddings_index.py."""

import numpy as np
import pytest

from oracle import search as osearch
from tests.helpers import GOLDEN, load_json, synth_cases, unit_docs


def test_unit_fixtures_match_reference_outputs():
    for case in load_json("search_unit.json"):
        got = osearch.find(case["metric"], case["limit"], np.array(case["query"]), unit_docs(case["order"]))
        assert [[d, c] for d, c, _ in got] == case["expected"], case


@pytest.mark.parametrize("metric", osearch.ALL_METRICS)
def test_reference_test_search_stability(metric):
    # reference tests/test_embeddings_index.py:25-49
    q = np.array([1.0, 0.0, 0.0])
    assert [(d, c) for d, c, _ in osearch.find(metric, 1, q, unit_docs("123"))] == [(0, 0)]
    assert [(d, c) for d, c, _ in osearch.find(metric, 1, q, unit_docs("321"))] == [(1, 0)]


@pytest.mark.parametrize("metric", osearch.ALL_METRICS)
@pytest.mark.parametrize("limit", [1, 2, 3, 10])
def test_reference_test_different_limits(metric, limit):
    # reference tests/test_embeddings_index.py:52-69
    got = osearch.find(metric, limit, np.array([1.0, 0.0, 0.0]), unit_docs("123"))
    assert [(d, c) for d, c, _ in got] == [(0, 0), (1, 0), (0, 1)][:limit]


@pytest.mark.parametrize("metric", osearch.ALL_METRICS)
def test_reference_test_empty_index(metric):
    # reference tests/test_embeddings_index.py:72-94
    q = np.array([0.0, 0.0, 0.0])
    assert osearch.find(metric, 1, q, []) == []
    assert osearch.find(metric, 1, q, unit_docs("3")) == []


def test_metric_known_answers_from_reference_tests():
    # reference tests/test_embeddings_metrics.py:6-201 (values restated)
    a = np.array
    e4 = a([[1.0, 0, 0, 0], [0, 1.0, 0, 0]])
    np.testing.assert_allclose(osearch.distances("cosine_sim", a([1.0, 0, 0, 0]), e4), [-1.0, 0.0])
    np.testing.assert_allclose(osearch.distances("cosine_sim", a([2.0, 0, 0, 0]), e4), [-1.0, 0.0])
    np.testing.assert_allclose(
        osearch.distances("cosine_sim", a([0.0, 0, 0, 0]), a([[1.0, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 0]])), [0, 0, 0])
    np.testing.assert_allclose(osearch.distances("inner_product", a([2, 0, 0, 0]), a([[1, 0, 0, 0], [0, 1, 0, 0]])), [-2.0, 0.0])
    np.testing.assert_allclose(
        osearch.distances("euclidean_dist", a([1, 0, 0, 0]), a([[2, 0, 0, 0], [3, 3, 3, 0], [0, 0, 0, 0]])),
        [1.0, np.sqrt(22), 1.0])
    np.testing.assert_allclose(
        osearch.distances("sqeuclidean_dist", a([0, 0, 0, 0]), a([[1, 1, 1, 1], [2, 2, 2, 2]])), [4.0, 16.0])
    q = a([1, 2, 3, 4])
    docs = a([[1, 0, 0, 0], [0, 1, 0, 0], [2, 0, 0, 0], [3, 3, 3, 0], [0, 0, 0, 0]])
    np.testing.assert_allclose(osearch.distances("euclidean_dist", q, docs) ** 2,
                               osearch.distances("sqeuclidean_dist", q, docs))


def test_metric_values_bitwise_equal_reference():
    kat = load_json("search_metrics_kat.json")
    docs = np.array(kat["docs"], dtype=np.float32)
    q = np.array(kat["query"], dtype=np.float64)
    for metric, hexes in kat["distances"].items():
        want = np.array([float.fromhex(h) for h in hexes])
        got = osearch.distances(metric, q, docs)
        assert np.array_equal(got, want, equal_nan=True), metric
    want = np.array([float.fromhex(h) for h in kat["zero_query_cosine"]])
    assert np.array_equal(osearch.distances("cosine_sim", np.zeros(16), docs), want)


def test_synth_cases_match_reference_outputs():
    n = 0
    for entry, data in synth_cases():
        for r in entry["results"]:
            got = osearch.find(r["metric"], r["limit"], data["queries"][r["query"]], data["docs"])
            assert [[d, c] for d, c, _ in got] == r["expected"], (entry["name"], r["metric"], r["limit"], r["query"])
            n += 1
        big = entry["in_doc"]["doc"]
        for row in entry["in_doc"]["rows"]:
            ids, dist = osearch.find_in_doc(row["metric"], 20, data["queries"][0], *data["docs"][big])
            assert ids.tolist() == row["chunk_ids"]
            want = np.array([float.fromhex(h) for h in row["distances"]])
            assert np.array_equal(dist, want, equal_nan=True)
    assert n >= 300


def test_small_inputs_npz_matches_generator():
    z = np.load(f"{GOLDEN}/search_small_inputs.npz")
    for entry, data in synth_cases():
        if entry["name"] != "small":
            continue
        assert np.array_equal(z["queries"], data["queries"])
        for i, (ids, emb) in enumerate(data["docs"]):
            assert np.array_equal(z[f"emb{i}"], emb) and np.array_equal(z[f"ids{i}"], ids)


def test_global_topk_equals_per_doc_then_global():
    # the identity the GPU path relies on
    for entry, data in synth_cases():
        if entry["name"] not in ("small", "odd_dim"):
            continue
        docs = [d for d in data["docs"] if len(d[1])]
        flat = np.concatenate([e for _, e in docs])
        offs = np.cumsum([0] + [len(e) for _, e in docs])
        for metric in osearch.ALL_METRICS:
            for q in data["queries"][:3]:
                rows, _ = osearch.topk_rows(metric, 9, q, flat)
                ref = osearch.find(metric, 9, q, docs)
                di = np.searchsorted(offs, rows, side="right") - 1
                got = [(int(d), int(docs[d][0][r - offs[d]])) for d, r in zip(di, rows)]
                assert got == [(d, c) for d, c, _ in ref]
