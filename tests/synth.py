"""Seeded synthetic inputs shared by the golden generator, the tests and bench.py.

Everything is derived from ``numpy.random.Generator(PCG64(seed))`` so that the
authoring container (which ran the reference to make ``tests/golden``) and the
GPU box rebuild identical arrays; the golden files carry sha256 digests of the
arrays to detect a drifting generator.
"""

from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def synth_matrix(seed: int, rows: int, dim: int, normalise: bool = True) -> np.ndarray:
    m = _rng(seed).standard_normal((rows, dim), dtype=np.float32)
    if normalise and rows:
        m /= np.linalg.norm(m, axis=1, keepdims=True)
    return m


def synth_queries(seed: int, n: int, dim: int, normalise: bool = True) -> np.ndarray:
    """float64 queries, as the reference builds them (semantic_retriever.py:49,53)."""
    q = _rng(seed).standard_normal((n, dim), dtype=np.float32)
    if normalise:
        q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(np.float64)


def synth_index(
    seed: int,
    doc_rows: Sequence[int],
    dim: int,
    n_queries: int,
    dup_rows: int = 0,
    multi_row_chunks: bool = False,
    normalise: bool = True,
) -> Dict:
    """A list of documents ``(chunk_ids int64[N_i], embeddings f32[N_i, dim])``.

    * ``dup_rows`` rows are overwritten with copies of other rows (possibly in
      another document) so that exact score ties exist and tie-breaking by
      document order / row order is exercised;
    * the first query is made an exact copy of a stored row (distance 0 / best
      score shared by all of its duplicates);
    * ``multi_row_chunks`` makes some chunks own several rows (repeated
      ``chunk_id``), as the description/multimodal indexes do
      (embeddings_index.py:101-136).
    """
    rng = _rng(seed)
    total = int(sum(doc_rows))
    flat = rng.standard_normal((total, dim), dtype=np.float32)
    if not normalise:
        flat *= rng.uniform(0.1, 3.0, size=(total, 1)).astype(np.float32)
    else:
        flat /= np.linalg.norm(flat, axis=1, keepdims=True)
    if total and dup_rows:
        src = rng.integers(0, total, size=dup_rows)
        dst = rng.integers(0, total, size=dup_rows)
        flat[dst] = flat[src]
    queries = rng.standard_normal((n_queries, dim), dtype=np.float32)
    if normalise:
        queries /= np.linalg.norm(queries, axis=1, keepdims=True)
    queries = queries.astype(np.float64)
    if total and n_queries:
        pick = int(rng.integers(0, total))
        queries[0] = flat[pick].astype(np.float64)
        if dup_rows:
            # make sure the exact-match row has duplicates elsewhere
            extra = rng.integers(0, total, size=3)
            flat[extra] = flat[pick]
    docs: List = []
    start = 0
    for n in doc_rows:
        emb = flat[start : start + n]
        start += n
        if n == 0:
            docs.append((np.array([], dtype=np.int64), np.array([], dtype=np.float32)))
            continue
        if multi_row_chunks:
            reps = rng.integers(1, 4, size=n)
            ids = np.repeat(np.arange(n), reps)[:n].astype(np.int64)
        else:
            ids = np.arange(n, dtype=np.int64)
        docs.append((ids, np.ascontiguousarray(emb)))
    return {"docs": docs, "queries": queries}


def synth_token_batch(seed: int, n_seq: int, seq_len: int, ragged: bool = False,
                      vocab: int = 30522, min_len: int = 16):
    """Token ids as SURVEY 8d describes: [CLS]=101 ... [SEP]=102, ids in [1000, vocab).

    Returns ``(ids int32[total], cu_seqlens int32[n_seq+1])`` (packed, no padding).
    """
    rng = _rng(seed)
    if ragged:
        lens = rng.integers(min_len, seq_len + 1, size=n_seq)
    else:
        lens = np.full(n_seq, seq_len, dtype=np.int64)
    cu = np.zeros(n_seq + 1, dtype=np.int32)
    cu[1:] = np.cumsum(lens)
    ids = rng.integers(1000, vocab, size=int(cu[-1]), dtype=np.int64).astype(np.int32)
    ids[cu[:-1]] = 101
    ids[cu[1:] - 1] = 102
    return ids, cu
