"""CPU ORACLE (test infrastructure, not product code) -- similarity search.

A numpy/torch restatement of the reference's brute-force search for the
semantic-retriever hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module; the product path (``ai-dial-rag_b200/``) never does.

Parity status: PINNED.  ``tests/golden/search_*.json`` were produced by running
the reference's *own* ``aidial_rag/retrievers/embeddings_metrics.py`` and
``embeddings_index.py`` in the authoring container (see
``oracle/make_golden_search.py``) and this restatement is checked against them
and against the known answers in the reference's
``tests/test_embeddings_metrics.py`` / ``tests/test_embeddings_index.py``.

Reference lines followed (relative to /root/reference):
  * metrics            aidial_rag/retrievers/embeddings_metrics.py:14-58
  * per-document top-k aidial_rag/retrievers/embeddings_index.py:51-60
  * cross-document     aidial_rag/retrievers/embeddings_index.py:62-89
  * index flattening   aidial_rag/retrievers/embeddings_index.py:101-136
"""

from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

COSINE_SIM = "cosine_sim"
EUCLIDEAN_DIST = "euclidean_dist"
SQEUCLIDEAN_DIST = "sqeuclidean_dist"
INNER_PRODUCT = "inner_product"
ALL_METRICS = (COSINE_SIM, EUCLIDEAN_DIST, SQEUCLIDEAN_DIST, INNER_PRODUCT)


def distances(metric: str, query: np.ndarray, docs: np.ndarray) -> np.ndarray:
    """"Smaller is better" distance of every row of ``docs`` to ``query``.

    embeddings_metrics.py:14-50.  The arithmetic (operation order, the dtypes
    each intermediate lives in) is kept as in the reference because the
    ranking is defined by these exact floating point values:

    * inner product (:20)  -> ``-(q . d)``; with the float64 query the
      reference passes (semantic_retriever.py:49,53) numpy promotes the float32
      matrix and accumulates in float64.
    * cosine (:28-31)      -> torch's ``cosine_similarity`` (eps-clamped norms,
      so all-zero vectors score 0 instead of NaN).
    * sq. euclid (:40-43)  -> ``sum(d*d) - 2*(d . q) + sum(q*q)``; ``sum(d*d)``
      is evaluated in the matrix dtype (float32, numpy pairwise summation)
      and only then promoted.
    * euclid (:50)         -> sqrt of the previous.
    """
    if metric == INNER_PRODUCT:
        return -np.inner(query, docs)
    if metric == COSINE_SIM:
        sim = torch.nn.functional.cosine_similarity(
            torch.from_numpy(docs), torch.from_numpy(query)
        )
        return -sim.numpy()
    if metric in (SQEUCLIDEAN_DIST, EUCLIDEAN_DIST):
        row_sq = np.sum(docs**2, axis=1)
        q_sq = np.sum(query**2)
        cross = np.dot(docs, query)
        sq = row_sq - 2 * cross + q_sq
        return np.sqrt(sq) if metric == EUCLIDEAN_DIST else sq
    raise ValueError(f"unknown metric {metric!r}")


def find_in_doc(
    metric: str,
    limit: int,
    query: np.ndarray,
    chunk_ids: np.ndarray,
    embeddings: np.ndarray,
) -> Tuple[np.ndarray, np.ndarray]:
    """embeddings_index.py:51-60 -- full stable argsort, then slice."""
    d = distances(metric, query, embeddings)
    order = np.argsort(d, kind="stable")[:limit]
    return chunk_ids[order], d[order]


def find(
    metric: str,
    limit: int,
    query: np.ndarray,
    doc_indexes: Sequence[Tuple[np.ndarray, np.ndarray]],
) -> List[Tuple[int, int, float]]:
    """embeddings_index.py:62-89.

    ``doc_indexes`` is a list of ``(chunk_ids int64[N_i], embeddings[N_i, D])``.
    Returns ``[(doc_id, chunk_id, distance), ...]`` best first.  Empty documents
    are skipped (:67-68); per-document winners are concatenated in document
    order and re-sorted with a second stable argsort (:81), so ties go to the
    earlier document, then to the lower row.
    """
    all_doc = np.array([], dtype=np.int64)
    all_chunk = np.array([], dtype=np.int64)
    all_dist = np.array([], dtype=np.float32)
    for doc_id, (chunk_ids, emb) in enumerate(doc_indexes):
        if len(emb) == 0:
            continue
        top_chunks, top_d = find_in_doc(metric, limit, query, chunk_ids, emb)
        all_doc = np.concatenate(
            (all_doc, np.full(len(top_chunks), doc_id, dtype=np.int64))
        )
        all_chunk = np.concatenate((all_chunk, top_chunks))
        all_dist = np.concatenate((all_dist, top_d))
    order = np.argsort(all_dist, kind="stable")[:limit]
    return [
        (int(all_doc[i]), int(all_chunk[i]), float(all_dist[i])) for i in order
    ]


def topk_rows(
    metric: str, limit: int, query: np.ndarray, matrix: np.ndarray
) -> Tuple[np.ndarray, np.ndarray]:
    """Global row ids + distances of the ``limit`` best rows of one matrix.

    Equivalent to ``find`` over the document-order concatenation (SURVEY 8a:
    per-doc-then-global stable selection == one global stable top-k).
    """
    if len(matrix) == 0:
        return np.array([], dtype=np.int64), np.array([], dtype=np.float64)
    d = distances(metric, query, matrix)
    order = np.argsort(d, kind="stable")[:limit]
    return order.astype(np.int64), d[order]


def flatten_by_chunk(items: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """embeddings_index.py:121-136: item i owns ``len(items[i])`` rows."""
    ids: List[int] = []
    rows: List[np.ndarray] = []
    for i, e in enumerate(items):
        ids.extend([i] * len(e))
        rows.extend(e)
    return np.array(ids, dtype=np.int64), np.array(rows)


def flatten_by_page(
    chunk_pages: Sequence[int], page_items: Sequence[np.ndarray]
) -> Tuple[np.ndarray, np.ndarray]:
    """embeddings_index.py:101-118: every chunk repeats its page's rows.

    ``chunk_pages[i]`` is the 0-based page of chunk ``i`` (:92-94).
    """
    ids: List[int] = []
    rows: List[np.ndarray] = []
    for i, page in enumerate(chunk_pages):
        e = page_items[page]
        ids.extend([i] * len(e))
        rows.extend(e)
    return np.array(ids, dtype=np.int64), np.array(rows, dtype=np.float32)
