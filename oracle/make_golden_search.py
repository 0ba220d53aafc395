"""Generate tests/golden/search_*.{json,npz} by running the REAL reference.

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden_search.py

It executes the reference's own ``EmbeddingsIndex.find`` / ``find_in_doc`` /
``ENUM_TO_METRIC`` (loaded through ``oracle/ref_shims.py``) on

  * the fixtures of the reference's tests/test_embeddings_index.py:11-94,
  * seeded synthetic multi-document indexes with ragged/empty documents,
    planted duplicate rows and multi-row chunks,

and stores inputs (npz, small) or their seed + sha256 (larger cases) next to the
reference's outputs.  ``tests/test_oracle_search.py`` pins ``oracle/search.py``
to these files; the GPU parity tests use the same files through the C-ABI.
"""

from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_shims import load_reference  # noqa: E402
from tests.synth import synth_index  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main() -> None:
    metrics_mod, index_mod, record_mod = load_reference()
    Metric = metrics_mod.Metric
    DocIndex, EmbeddingsIndex = index_mod.DocIndex, index_mod.EmbeddingsIndex
    TEXT = record_mod.RetrievalType.TEXT
    os.makedirs(GOLDEN, exist_ok=True)

    def run_find(docs, query, metric, limit):
        index = EmbeddingsIndex(
            retrieval_type=TEXT,
            indexes=[DocIndex(chunk_ids=c, embeddings=e) for c, e in docs],
            metric=Metric(metric),
            limit=limit,
        )
        out = index.find(query)
        return [[int(d.metadata["doc_id"]), int(d.metadata["chunk_id"])] for d in out]

    def run_find_in_doc(doc, query, metric, limit):
        index = EmbeddingsIndex(
            retrieval_type=TEXT, indexes=[], metric=Metric(metric), limit=limit
        )
        ids, dist = index.find_in_doc(query, DocIndex(chunk_ids=doc[0], embeddings=doc[1]))
        return [int(i) for i in ids], [float(x).hex() for x in dist]

    # ---- 1. the reference's own unit fixtures (test_embeddings_index.py:11-22)
    doc1 = (np.array([0, 1], dtype=np.int64),
            np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], dtype=np.float32))
    doc2 = (np.array([0], dtype=np.int64),
            np.array([[1.0, 0.0, 0.0]], dtype=np.float32))
    doc3 = (np.array([], dtype=np.int64), np.array([], dtype=np.float32))
    unit = []
    q = np.array([1.0, 0.0, 0.0])
    for metric in Metric:
        for limit in (1, 2, 3, 10):
            for name, docs in (("123", [doc1, doc2, doc3]), ("321", [doc3, doc2, doc1])):
                unit.append({"metric": str(metric), "limit": limit, "order": name,
                             "query": q.tolist(),
                             "expected": run_find(docs, q, metric, limit)})
        z = np.array([0.0, 0.0, 0.0])
        unit.append({"metric": str(metric), "limit": 1, "order": "empty", "query": z.tolist(),
                     "expected": run_find([], z, metric, 1)})
        unit.append({"metric": str(metric), "limit": 1, "order": "3", "query": z.tolist(),
                     "expected": run_find([doc3], z, metric, 1)})
    with open(os.path.join(GOLDEN, "search_unit.json"), "w") as f:
        json.dump(unit, f, indent=1)

    # ---- 2. metric known-answer vectors straight from the reference functions
    rng = np.random.default_rng(11)
    kat_docs = rng.standard_normal((64, 16)).astype(np.float32)
    kat_docs[7] = 0.0  # zero row: cosine must give 0, not NaN
    kat_docs[9] = kat_docs[3]
    kat_q = rng.standard_normal(16)
    kat = {"docs": kat_docs.tolist(), "query": kat_q.tolist(), "distances": {}}
    for metric in Metric:
        d = metrics_mod.ENUM_TO_METRIC[metric](kat_q, kat_docs)
        kat["distances"][str(metric)] = [float(x).hex() for x in d]
    zq = np.zeros(16)
    kat["zero_query_cosine"] = [
        float(x).hex() for x in metrics_mod.ENUM_TO_METRIC[Metric.COSINE_SIM](zq, kat_docs)
    ]
    with open(os.path.join(GOLDEN, "search_metrics_kat.json"), "w") as f:
        json.dump(kat, f)

    # ---- 3. seeded synthetic multi-document indexes
    cases = []
    small = synth_index(seed=101, doc_rows=[57, 0, 1, 130, 3, 0, 109], dim=384,
                        n_queries=6, dup_rows=12, multi_row_chunks=True)
    np.savez_compressed(
        os.path.join(GOLDEN, "search_small_inputs.npz"),
        queries=small["queries"],
        **{f"emb{i}": e for i, (_, e) in enumerate(small["docs"])},
        **{f"ids{i}": c for i, (c, _) in enumerate(small["docs"])},
    )
    specs = [
        ("small", dict(seed=101, doc_rows=[57, 0, 1, 130, 3, 0, 109], dim=384,
                       n_queries=6, dup_rows=12, multi_row_chunks=True)),
        ("medium", dict(seed=202, doc_rows=[4000, 1, 0, 2500, 777, 3000, 9], dim=384,
                        n_queries=8, dup_rows=200, multi_row_chunks=False)),
        ("odd_dim", dict(seed=303, doc_rows=[300, 45, 1000], dim=100,
                         n_queries=4, dup_rows=20, multi_row_chunks=True)),
        ("wide_dim", dict(seed=404, doc_rows=[500, 300], dim=1024,
                          n_queries=3, dup_rows=10, multi_row_chunks=False)),
        ("unnormalised", dict(seed=505, doc_rows=[800, 800], dim=384, n_queries=4,
                              dup_rows=30, multi_row_chunks=False, normalise=False)),
    ]
    for name, spec in specs:
        data = synth_index(**spec)
        entry = {"name": name, "spec": spec,
                 "sha_matrix": sha(np.concatenate([e.reshape(-1, spec["dim"]) for _, e in data["docs"]])),
                 "sha_queries": sha(data["queries"]), "results": []}
        for metric in Metric:
            for limit in (1, 7, 100):
                for qi, query in enumerate(data["queries"]):
                    entry["results"].append({
                        "metric": str(metric), "limit": limit, "query": qi,
                        "expected": run_find(data["docs"], query, metric, limit)})
        # find_in_doc distances of the biggest document, limit 20 (hex floats)
        big = max(range(len(data["docs"])), key=lambda i: len(data["docs"][i][1]))
        entry["in_doc"] = {"doc": big, "rows": []}
        for metric in Metric:
            ids, dist = run_find_in_doc(data["docs"][big], data["queries"][0], metric, 20)
            entry["in_doc"]["rows"].append({"metric": str(metric), "chunk_ids": ids, "distances": dist})
        cases.append(entry)
    with open(os.path.join(GOLDEN, "search_synth.json"), "w") as f:
        json.dump(cases, f)
    print("wrote", os.listdir(GOLDEN))


if __name__ == "__main__":
    main()
