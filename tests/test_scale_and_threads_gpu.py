"""GPU tests at BASELINE sizes and under the reference's threading contract.

* configs[2]: exact top-100 over a 10M x 384 fp32 index, ids against a blocked numpy restatement of the reference path
  (float64 query, ``-(docs . q)``, stable ascending order, aidial_rag/retrievers/embeddings_metrics.py:14-20 +
  embeddings_index.py:57-59) -- through the batched tensor-core path AND the float64 scan;
* limit > MAX_K (the reference's ``argsort()[:limit]`` takes any limit);
* SURVEY 8b threading: two threads inside the encoder at once (the indexing and the query pool, cpu_pools.py:50-59)
  and ``find`` from many executor threads (semantic_retriever.py:54-56).
"""

import threading

import numpy as np
import pytest
import torch

from oracle import encoder as oenc
from oracle import search as osearch
from tests.synth import synth_matrix, synth_queries, synth_token_batch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def test_config3_full_size_ids_vs_blocked_numpy_oracle():
    from dial_rag_b200.device_index import DeviceMatrix

    rows, dim, k, nq = 10_000_000, 384, 100, 4
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(2)
    mat = torch.empty((rows, dim), dtype=torch.float32, device=dev)
    blk = 1 << 20
    for s in range(0, rows, blk):
        x = torch.randn((min(blk, rows - s), dim), generator=g, device=dev)
        mat[s:s + blk] = x / x.norm(dim=1, keepdim=True)
    # planted duplicates: ties must resolve to the lowest row id
    mat[9_000_001] = mat[17]
    mat[5_000_000] = mat[17]
    q = synth_queries(seed=3, n=nq, dim=dim)
    q[0] = mat[17].double().cpu().numpy()          # exact match of a triplicated row
    # oracle, block by block on the host: best k of every block, then best k of the union (stable by row id)
    cand_d, cand_r = [[] for _ in range(nq)], [[] for _ in range(nq)]
    for s in range(0, rows, blk):
        docs = mat[s:s + blk].cpu().numpy()
        for i in range(nq):
            r, d = osearch.topk_rows("inner_product", k, q[i], docs)
            cand_r[i].append(r + s)
            cand_d[i].append(d)
    want = []
    for i in range(nq):
        r, d = np.concatenate(cand_r[i]), np.concatenate(cand_d[i])
        order = np.lexsort((r, d))[:k]              # ascending distance, ties by row id
        want.append(r[order])
    dm = DeviceMatrix(mat)
    del mat
    _, got_batch, _ = dm.topk(q, k, "inner_product")
    assert dm._use_batch(nq, k, 3) and dm.last_batch_fallbacks == 0
    dq = torch.from_numpy(q).to(dev)
    got_scan = dm.topk_device(dq, k, "inner_product", allow_batch=False)[1].cpu().numpy()
    for i in range(nq):
        assert np.array_equal(got_batch[i], want[i]), (i, "batch")
        assert np.array_equal(got_scan[i], want[i]), (i, "scan")
    assert got_batch[0][:3].tolist() == [17, 5_000_000, 9_000_001]


@pytest.mark.parametrize("metric", ["sqeuclidean_dist", "cosine_sim"])
def test_limit_above_max_k(metric):
    from dial_rag_b200._native import MAX_K
    from dial_rag_b200.device_index import DeviceMatrix

    m = synth_matrix(seed=8, rows=6000, dim=96)
    m[4000] = m[10]
    q = synth_queries(seed=9, n=2, dim=96)
    k = MAX_K + 1500
    dm = DeviceMatrix(m)
    dist, rows, count = dm.topk(q, k, metric)
    assert rows.shape == (2, k) and (count == k).all()
    for i in range(2):
        want_rows, want_dist = osearch.topk_rows(metric, k, q[i], m)
        assert np.array_equal(rows[i], want_rows)
        np.testing.assert_allclose(dist[i], want_dist, rtol=1e-9, atol=3e-7)
    all_rows = dm.topk(q[:1], 10_000, metric)[1]     # limit > N returns all N
    assert all_rows.shape == (1, 6000) and sorted(all_rows[0].tolist()) == list(range(6000))


def test_indexing_and_query_threads_inside_the_encoder_at_once():
    """One thread streams indexing batches (bulk workspace), another embeds single queries (query workspace, own
    high-priority stream) at the same time; every result equals the one computed alone."""
    from dial_rag_b200.embeddings.encoder import B200Encoder

    w = oenc.synth_weights(seed=7, style="stress")
    enc = B200Encoder(w, device=0, max_tokens=65536)
    try:
        big = [synth_token_batch(seed=50 + i, n_seq=200, seq_len=256, ragged=True) for i in range(3)]
        small = [synth_token_batch(seed=80 + i, n_seq=1, seq_len=300 + 40 * i) for i in range(5)]
        want_big = [enc.embed_packed(*b) for b in big]
        want_small = [enc.embed_packed(*s) for s in small]
        errors, done = [], threading.Event()

        def indexing():
            try:
                for rep in range(6):
                    for b, wb in zip(big, want_big):
                        assert np.array_equal(enc.embed_packed(*b), wb)
            except BaseException as e:  # noqa: BLE001
                errors.append(e)
            finally:
                done.set()

        def querying():
            try:
                n = 0
                while not done.is_set() or n < 20:
                    i = n % len(small)
                    assert np.array_equal(enc.embed_packed(*small[i]), want_small[i])
                    n += 1
            except BaseException as e:  # noqa: BLE001
                errors.append(e)

        threads = [threading.Thread(target=indexing), threading.Thread(target=querying)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=600)
        assert not errors, errors
        assert not any(t.is_alive() for t in threads)
    finally:
        enc.close()


def test_device_forwards_on_different_streams_share_the_workspace_safely():
    """two asynchronous forwards on different streams must not run concurrently over one workspace."""
    from dial_rag_b200.embeddings.encoder import B200Encoder

    w = oenc.synth_weights(seed=0, style="hf_init")
    enc = B200Encoder(w, device=0, max_tokens=65536)
    dev = torch.device("cuda", 0)
    try:
        batches = [synth_token_batch(seed=60 + i, n_seq=160, seq_len=256) for i in range(4)]
        want = [enc.embed_packed(*b) for b in batches]
        streams = [torch.cuda.Stream(dev) for _ in range(4)]
        outs = []
        torch.cuda.synchronize(dev)
        for (ids, cu), st in zip(batches, streams):
            d_ids, d_cu = torch.from_numpy(ids).to(dev), torch.from_numpy(cu).to(dev)
            out = torch.empty((len(cu) - 1, 384), dtype=torch.float32, device=dev)
            torch.cuda.synchronize(dev)
            outs.append((d_ids, d_cu, out))
        for rep in range(3):
            for (ids, cu), st, (d_ids, d_cu, out) in zip(batches, streams, outs):
                enc.forward_device(d_ids, d_cu, cu, out, stream=st.cuda_stream)   # back to back, no host sync in between
        torch.cuda.synchronize(dev)
        for (_, _, out), wb in zip(outs, want):
            assert np.array_equal(out.cpu().numpy(), wb)
    finally:
        enc.close()


def test_find_from_eight_threads():
    from dial_rag_b200.records import RetrievalType
    from dial_rag_b200.retrievers.embeddings_index import DocIndex, EmbeddingsIndex

    docs = [DocIndex(np.arange(n, dtype=np.int64), synth_matrix(seed=20 + i, rows=n, dim=384)) for i, n in enumerate((30_000, 1, 12_000))]
    index = EmbeddingsIndex(RetrievalType.TEXT, docs, limit=7)
    queries = synth_queries(seed=33, n=24, dim=384)
    want = [index.find(q) for q in queries]
    batch = index.find_batch(queries)
    assert batch == want
    errors = []

    def worker(t):
        try:
            for rep in range(10):
                for i in range(t, len(queries), 8):
                    assert index.find(queries[i]) == want[i]
        except BaseException as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
