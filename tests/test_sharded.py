"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes,
oracle plugged in as the local searcher / merger (no CUDA in this container)."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_local_search(rows, row_start, queries, k, metric):
    from oracle import search as osearch

    nq = len(queries)
    d = np.full((nq, k), np.nan)
    r = np.full((nq, k), -1, dtype=np.int64)
    c = np.zeros(nq, dtype=np.int32)
    for i in range(nq):
        if len(rows) == 0:
            continue
        ids, dd = osearch.topk_rows(metric, k, queries[i], rows)
        d[i, : len(ids)] = dd
        r[i, : len(ids)] = ids + row_start
        c[i] = len(ids)
    return d, r, c


def _oracle_merge(gd, gr, gc, k):
    """Stable merge by (distance, NaN last, row id) -- what drag_topk_merge implements."""
    s, nq, _ = gd.shape
    md = np.full((nq, k), np.nan)
    mr = np.full((nq, k), -1, dtype=np.int64)
    mc = np.zeros(nq, dtype=np.int32)
    for q in range(nq):
        cand = [(gd[i, q, j], gr[i, q, j]) for i in range(s) for j in range(gc[i, q])]
        cand.sort(key=lambda t: (np.isnan(t[0]), t[0] if not np.isnan(t[0]) else 0.0, t[1]))
        cand = cand[:k]
        mc[q] = len(cand)
        for j, (dd, rr) in enumerate(cand):
            md[q, j], mr[q, j] = dd, rr
    return md, mr, mc


def _worker(rank, world, port, n_rows, k, metric, out_dir):
    for p in (ROOT, os.path.join(ROOT, "ai-dial-rag_b200")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dial_rag_b200.sharded import ShardedIndex, split_contiguous
    from tests.synth import synth_matrix, synth_queries

    m = synth_matrix(seed=21, rows=n_rows, dim=64)
    if n_rows > 40:
        m[n_rows - 3] = m[1]  # tie across shards -> lower global row id must win
    q = synth_queries(seed=22, n=5, dim=64)
    if n_rows > 2:
        q[0] = m[1].astype(np.float64)
    start, end = split_contiguous(n_rows, world)[rank]
    idx = ShardedIndex(m[start:end], start, local_search=_oracle_local_search, merge=_oracle_merge)
    d, r, c = idx.topk(q, k, metric)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), d=d, r=r, c=c)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows,k", [(1000, 10), (7, 10), (1, 3), (64, 64)])
@pytest.mark.parametrize("metric", ["inner_product", "sqeuclidean_dist"])
def test_sharded_topk_two_ranks_gloo(tmp_path, n_rows, k, metric):
    from oracle import search as osearch
    from tests.synth import synth_matrix, synth_queries

    port = 29500 + (os.getpid() + n_rows + k) % 2000
    mp.spawn(_worker, args=(2, port, n_rows, k, metric, str(tmp_path)), nprocs=2, join=True)
    m = synth_matrix(seed=21, rows=n_rows, dim=64)
    if n_rows > 40:
        m[n_rows - 3] = m[1]
    q = synth_queries(seed=22, n=5, dim=64)
    if n_rows > 2:
        q[0] = m[1].astype(np.float64)
    res = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    for key in ("d", "r", "c"):
        assert np.array_equal(res[0][key], res[1][key], equal_nan=True), "ranks disagree"
    for i in range(len(q)):
        want_rows, want_d = osearch.topk_rows(metric, k, q[i], m)
        n = len(want_rows)
        assert res[0]["c"][i] == n
        assert np.array_equal(res[0]["r"][i, :n], want_rows)
        assert np.array_equal(res[0]["d"][i, :n], want_d)
        assert np.all(res[0]["r"][i, n:] == -1)


def test_split_contiguous_properties():
    from dial_rag_b200.sharded import split_contiguous

    for n in (0, 1, 7, 8, 1000, 100_000_000):
        for w in (1, 2, 3, 8):
            b = split_contiguous(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and len(b) == w
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    from dial_rag_b200.sharded import pack_candidates, unpack_candidates

    d = torch.tensor([[0.5, float("nan")], [-1.25, 3.0]], dtype=torch.float64)
    r = torch.tensor([[7, -1], [2, 9]], dtype=torch.int64)
    c = torch.tensor([1, 2], dtype=torch.int32)
    buf = pack_candidates(torch, d, r, c)
    d2, r2, c2 = unpack_candidates(torch, buf[None], 2)
    assert torch.equal(r2[0], r) and torch.equal(c2[0], c)
    assert np.array_equal(d2[0].numpy(), d.numpy(), equal_nan=True)
