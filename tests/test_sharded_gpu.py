"""CUDA side of the row-sharded index: single-rank path always; 2-rank NCCL path when the box has
two GPUs (gpurun --gpus 2)."""

import os
import sys

import numpy as np
import pytest
import torch

from oracle import search as osearch
from tests.synth import synth_matrix, synth_queries

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_single_rank_sharded_index_matches_oracle():
    from dial_rag_b200.sharded import ShardedIndex

    m = synth_matrix(seed=31, rows=30_000, dim=384)
    q = synth_queries(seed=32, n=3, dim=384)
    idx = ShardedIndex(m, row_start=1000)
    d, r, c = idx.topk(q, 20, "sqeuclidean_dist")
    for i in range(len(q)):
        want_rows, want_d = osearch.topk_rows("sqeuclidean_dist", 20, q[i], m)
        assert np.array_equal(r[i], want_rows + 1000)
        np.testing.assert_allclose(d[i], want_d, rtol=1e-9, atol=1e-9)
    assert c.tolist() == [20, 20, 20]


def _nccl_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "ai-dial-rag_b200")):
        sys.path.insert(0, p)
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from dial_rag_b200.sharded import ShardedIndex, split_contiguous

    m = synth_matrix(seed=41, rows=100_001, dim=384)
    m[90_000] = m[5]
    q = synth_queries(seed=42, n=9, dim=384)
    q[0] = m[5].astype(np.float64)
    start, end = split_contiguous(len(m), world)[rank]
    for storage in ("f32", "bf16"):
        idx = ShardedIndex(m[start:end], start, storage=storage, device=rank)
        d, r, c = idx.topk(q, 100, "inner_product")
        np.savez(os.path.join(out_dir, f"{storage}_rank{rank}.npz"), d=d, r=r, c=c)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_allgather_merge(tmp_path):
    import torch.multiprocessing as mp

    mp.spawn(_nccl_worker, args=(2, 29611, str(tmp_path)), nprocs=2, join=True)
    m = synth_matrix(seed=41, rows=100_001, dim=384)
    m[90_000] = m[5]
    q = synth_queries(seed=42, n=9, dim=384)
    q[0] = m[5].astype(np.float64)
    for storage in ("f32", "bf16"):
        ref = m if storage == "f32" else torch.from_numpy(m).to(torch.bfloat16).float().numpy()
        a, b = (np.load(tmp_path / f"{storage}_rank{r}.npz") for r in range(2))
        assert np.array_equal(a["r"], b["r"]) and np.array_equal(a["d"], b["d"])
        for i in range(len(q)):
            want_rows, want_d = osearch.topk_rows("inner_product", 100, q[i], ref)
            assert np.array_equal(a["r"][i], want_rows), (storage, i)
            np.testing.assert_allclose(a["d"][i], want_d, rtol=1e-9, atol=1e-9)
