"""Multi-GPU layer of the hot path: one process per GPU, torch.distributed.

* Embedding is data parallel: weights are replicated, every rank embeds a contiguous slice of
  the chunk stream, no collective in the forward (``split_contiguous`` + ``embed_slice``).
* The index is ROW-SHARDED: rank r holds rows ``[start_r, end_r)`` of the global matrix and
  scans only those.  Queries are replicated; every rank produces its local exact top-k
  ``(distance f64, global row id i64)``, ONE all-gather exchanges the candidates (Q*(2k+1)*8
  bytes per rank) and every rank merges them with ``drag_topk_merge``.  The result is exact for
  any shard count because the global top-k is a subset of the union of the local top-ks, and
  global row ids keep the tie-break (lowest row id) identical to the single-GPU scan.

The local search and the merge are injectable so that the host-side logic (partitioning,
packing, gather, empty shards) is testable on CPU with the gloo backend (tests/test_sharded.py
plugs the oracle in); the defaults call the CUDA library.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def split_contiguous(n_items: int, world: int) -> List[Tuple[int, int]]:
    """Balanced contiguous partition: rank r owns [start, end); sizes differ by at most 1."""
    base, extra = divmod(int(n_items), int(world))
    bounds, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        bounds.append((start, start + size))
        start += size
    return bounds


def pack_candidates(torch, dist, rows, count):
    """[Q,k] f64, [Q,k] i64, [Q] i32 -> one int64 [Q, 2k+1] buffer (single collective)."""
    q, k = dist.shape
    buf = torch.empty((q, 2 * k + 1), dtype=torch.int64, device=dist.device)
    buf[:, :k] = dist.contiguous().view(torch.int64)
    buf[:, k:2 * k] = rows
    buf[:, 2 * k] = count.to(torch.int64)
    return buf


def unpack_candidates(torch, buf, k: int):
    """int64 [S, Q, 2k+1] -> (dist f64 [S,Q,k], rows i64 [S,Q,k], count i32 [S,Q])."""
    dist = buf[..., :k].contiguous().view(torch.float64)
    rows = buf[..., k:2 * k].contiguous()
    count = buf[..., 2 * k].to(torch.int32).contiguous()
    return dist, rows, count


class ShardedIndex:
    """Row-sharded exact top-k over the ranks of a process group."""

    def __init__(
        self,
        local_rows: np.ndarray,
        row_start: int,
        group=None,
        storage: str = "f32",
        device: Optional[int] = None,
        local_search: Optional[Callable] = None,
        merge: Optional[Callable] = None,
    ):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.row_start = int(row_start)
        self.n_local = int(len(local_rows))
        self.dim = int(local_rows.shape[1]) if self.n_local else None
        self._local_search = local_search
        self._merge = merge
        self._matrix = None
        if local_search is None:
            from dial_rag_b200.device_index import DeviceMatrix

            if self.n_local:
                self._matrix = DeviceMatrix(local_rows, device=device, storage=storage, row_id_base=self.row_start)
            self.comm_device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        else:
            self.comm_device = torch.device("cpu")
            self._rows = local_rows

    # ---- local stage -------------------------------------------------------------
    def _search_local(self, queries: np.ndarray, k: int, metric):
        torch = self.torch
        nq = len(queries)
        if self._local_search is not None:
            d, r, c = self._local_search(self._rows, self.row_start, queries, k, metric)
            return (torch.from_numpy(np.ascontiguousarray(d)), torch.from_numpy(np.ascontiguousarray(r)),
                    torch.from_numpy(np.ascontiguousarray(c)))
        if self._matrix is None:  # empty shard: nothing to offer
            return (torch.full((nq, k), float("nan"), dtype=torch.float64, device=self.comm_device),
                    torch.full((nq, k), -1, dtype=torch.int64, device=self.comm_device),
                    torch.zeros((nq,), dtype=torch.int32, device=self.comm_device))
        dq = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float64)).to(self.comm_device, non_blocking=True)
        k_local = min(k, self._matrix.n_rows)
        d, r, c = self._matrix.topk_device(dq, k_local, metric)
        if k_local < k:  # pad to the common width
            pad = k - k_local
            d = torch.cat((d, torch.full((nq, pad), float("nan"), dtype=torch.float64, device=d.device)), 1)
            r = torch.cat((r, torch.full((nq, pad), -1, dtype=torch.int64, device=r.device)), 1)
        return d, r, c

    # ---- exchange + merge --------------------------------------------------------
    def topk(self, queries: np.ndarray, k: int, metric, timers: Optional[dict] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Replicated queries in, identical global ``(dist, rows, count)`` out on every rank.
        ``timers`` (a dict, CUDA path only) receives the milliseconds of the call's stages -- local search (query
        upload included), candidate packing, all-gather, merge, download -- from CUDA events on the current stream."""
        torch = self.torch
        queries = np.asarray(queries, dtype=np.float64)
        if queries.ndim == 1:
            queries = queries[None, :]
        marks = []

        def mark(name):
            if timers is not None and self.comm_device.type == "cuda":
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(torch.cuda.current_stream(self.comm_device))
                marks.append((name, ev))

        mark("start")
        d, r, c = self._search_local(queries, k, metric)
        mark("local_search")
        packed = pack_candidates(torch, d, r, c)
        mark("pack")
        if self.world > 1:
            flat = torch.empty((self.world * packed.shape[0], packed.shape[1]), dtype=torch.int64, device=packed.device)
            self.dist.all_gather_into_tensor(flat, packed, group=self.group)  # ncclAllGather on GPUs
            gathered = flat.view(self.world, packed.shape[0], packed.shape[1])
        else:
            gathered = packed[None]
        mark("all_gather")
        gd, gr, gc = unpack_candidates(torch, gathered, k)
        if self._merge is not None:
            md, mr, mc = self._merge(gd.numpy(), gr.numpy(), gc.numpy(), k)
            return md, mr, mc
        from dial_rag_b200 import _native
        from dial_rag_b200.device_index import merge_topk_device

        md, mr, mc = merge_topk_device(_native.load(), self.comm_device.index, gd, gr, gc, k)
        mark("merge")
        out = md.cpu().numpy(), mr.cpu().numpy(), mc.cpu().numpy()
        mark("download")
        if marks:
            torch.cuda.synchronize(self.comm_device)
            for (_, a), (name, b) in zip(marks, marks[1:]):
                timers[name] = timers.get(name, 0.0) + a.elapsed_time(b)
        return out


def embed_slice(encoder, token_lists: Sequence[Sequence[int]], rank: int, world: int) -> Tuple[np.ndarray, Tuple[int, int]]:
    """Data-parallel embedding: this rank's contiguous slice of the chunk stream -> ``[n_r, 384]``
    float32 rows that land directly in its index shard (no collective)."""
    start, end = split_contiguous(len(token_lists), world)[rank]
    return encoder.embed_token_lists(token_lists[start:end]), (start, end)
