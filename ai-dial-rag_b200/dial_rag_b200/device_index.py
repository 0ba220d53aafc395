"""Device-resident chunk-embedding matrix + the calls into drag_topk.

PyTorch is used only as the owner of device/pinned buffers and of the CUDA stream
(``tensor.data_ptr()`` / ``torch.cuda.current_stream().cuda_stream`` are what crosses
the C ABI); every number is produced by libdrag_b200.so.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from dial_rag_b200 import _native
from dial_rag_b200._native import (BATCH_MAX_DIM, BATCH_MAX_K, DTYPE_BF16, DTYPE_F32, MAX_K, METRIC_CODES,
                                   DragError)


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise DragError(-2, "no CUDA device: dial_rag_b200 has no CPU path (it needs an sm_100 GPU)")
    return torch


def _metric_code(metric) -> int:
    try:
        return METRIC_CODES[str(getattr(metric, "value", metric))]
    except KeyError:
        raise ValueError(f"unknown metric {metric!r}") from None


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


class DeviceMatrix:
    """A row-major ``[n_rows, dim]`` embedding matrix resident in HBM.

    ``row_id_base`` is this shard's offset in the global row numbering (row-sharded
    index across GPUs); ``storage`` is ``"f32"`` (bit-exact path) or ``"bf16"``.
    """

    def __init__(
        self,
        matrix,
        device: Optional[int] = None,
        storage: str = "f32",
        row_id_base: int = 0,
        chunk_ids: Optional[np.ndarray] = None,
        doc_offsets: Optional[Sequence[int]] = None,
    ):
        torch = _torch()
        self.lib = _native.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        dev = torch.device("cuda", self.device)
        if isinstance(matrix, np.ndarray):
            if matrix.ndim != 2:
                raise ValueError(f"matrix must be 2-d, got {matrix.shape}")
            host = np.ascontiguousarray(matrix, dtype=np.float32)
            mat = torch.from_numpy(host).to(dev)
        else:  # a torch tensor already on the device (index built on the GPU)
            mat = matrix.to(dev)
            if mat.dim() != 2:
                raise ValueError(f"matrix must be 2-d, got {tuple(mat.shape)}")
        if storage == "bf16":
            mat = mat.to(torch.bfloat16)
            self.dtype_code = DTYPE_BF16
        elif storage == "f32":
            mat = mat.to(torch.float32)
            self.dtype_code = DTYPE_F32
        else:
            raise ValueError(f"storage must be 'f32' or 'bf16', got {storage!r}")
        self.matrix = mat.contiguous()
        self.n_rows, self.dim = int(mat.shape[0]), int(mat.shape[1])
        self.row_id_base = int(row_id_base)
        self.row_sq = torch.empty(self.n_rows, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _native.check(
                self.lib.drag_row_sqnorm(_ptr(self.matrix), self.dtype_code, self.n_rows, self.dim,
                                         _ptr(self.row_sq), stream)
            )
        self.chunk_ids = None
        if chunk_ids is not None:
            self.chunk_ids = torch.from_numpy(np.ascontiguousarray(chunk_ids, dtype=np.int64)).to(dev)
        offs = [0, self.n_rows] if doc_offsets is None else list(doc_offsets)
        self.n_docs = len(offs) - 1
        self.doc_offsets = torch.tensor(offs, dtype=torch.int64, device=dev)
        self._ws = {}
        # batched tensor-core path (built lazily on the first query batch that qualifies).  Crossover measured on
        # B200, 384-d fp32, top-20 (scripts/batch_threshold_probe.py): over 1M rows the batch path takes 0.35 ms
        # for any 1..16 queries against 0.55 / 1.10 / 2.19 ms for the float64 scan at 2 / 4 / 8 queries; over 100k
        # rows 0.165 ms against 0.156 / 0.242 / 0.470 ms.  None = pick by index size; an int overrides.
        self.batch_min_queries = None
        self.batch_min_rows = 8192
        self._batch_state = None
        self.last_batch_fallbacks = 0

    def nbytes(self) -> int:
        """Device bytes this index pins right now (rows, squared norms, id maps and -- once a query batch has built
        it -- the bf16 scoring copy and inverse norms of the batched path)."""
        total = self.matrix.numel() * self.matrix.element_size() + self.row_sq.numel() * 4 + self.doc_offsets.numel() * 8
        if self.chunk_ids is not None:
            total += self.chunk_ids.numel() * 8
        state = self._batch_state
        if state and state.get("ok") is not None and "shadow" in state:
            if state["shadow"] is not self.matrix:
                total += state["shadow"].numel() * 2
            total += state["inv"].numel() * 4
        return int(total)

    # ------------------------------------------------------------------ helpers
    def _check_queries(self, queries: np.ndarray) -> np.ndarray:
        q = np.asarray(queries)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(
                f"shapes not aligned: query has dimension {q.shape[-1]}, index rows have {self.dim}"
            )
        # the reference always scores with a float64 query (semantic_retriever.py:49,53)
        return np.ascontiguousarray(q, dtype=np.float64)

    def _workspace(self, torch, dev, n_queries: int, k: int):
        need = C.c_size_t(0)
        _native.check(self.lib.drag_topk_workspace_bytes(self.device, n_queries, k, C.byref(need)))
        return torch.empty(max(need.value, 16), dtype=torch.uint8, device=dev), need.value

    # ------------------------------------------------------------------ batched path
    def _batch_prepare(self):
        """bf16 scoring copy of the rows, inverse norms and norm statistics (once per index)."""
        if self._batch_state is not None:
            return self._batch_state
        torch = _torch()
        dev = self.matrix.device
        state = {"ok": False, "cos_ok": False}   # wide rows (D > 512: description / multimodal retrievers): float64 scan only
        if self.dim % 64 == 0 and self.dim <= BATCH_MAX_DIM and 0 < self.n_rows < 2 ** 31:
            stream = torch.cuda.current_stream(dev).cuda_stream
            inv = torch.empty(self.n_rows, dtype=torch.float32, device=dev)
            stats = torch.empty(4, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _native.check(self.lib.drag_row_norm_stats(_ptr(self.row_sq), self.n_rows, _ptr(inv), _ptr(stats), stream))
                if self.dtype_code == DTYPE_F32:
                    shadow = torch.empty((self.n_rows, self.dim), dtype=torch.bfloat16, device=dev)
                    _native.check(self.lib.drag_rows_to_bf16(_ptr(self.matrix), self.n_rows * self.dim, _ptr(shadow), stream))
                else:
                    shadow = self.matrix
            max_sq, min_sq, bad, _ = (float(x) for x in stats.cpu())
            # the error certificate needs finite rows of sane magnitude (else: float64 scan only)
            state = {
                "ok": bad == 0 and max_sq <= 1e30,
                "cos_ok": bad == 0 and max_sq <= 1e30 and (min_sq >= 1e-24 or min_sq == float("inf")),
                "shadow": shadow, "inv": inv, "max_norm": float(np.sqrt(max_sq)),
            }
        self._batch_state = state
        return state

    def _batch_threshold(self) -> int:
        if self.batch_min_queries is not None:
            return int(self.batch_min_queries)
        return 2 if self.n_rows >= 262144 else (3 if self.n_rows >= 65536 else 8)

    def _use_batch(self, nq: int, k: int, metric_code: int) -> bool:
        if nq < self._batch_threshold() or self.n_rows < self.batch_min_rows or k > BATCH_MAX_K:
            return False
        state = self._batch_prepare()
        return bool(state["cos_ok"] if metric_code == METRIC_CODES["cosine_sim"] else state["ok"])

    def _topk_batch_device(self, d_queries, k: int, metric_code: int):
        torch = _torch()
        dev = self.matrix.device
        state = self._batch_state
        nq = int(d_queries.shape[0])
        dist = torch.empty((nq, k), dtype=torch.float64, device=dev)
        rows = torch.empty((nq, k), dtype=torch.int64, device=dev)
        count = torch.empty((nq,), dtype=torch.int32, device=dev)
        status = torch.empty((nq,), dtype=torch.int32, device=dev)
        need = C.c_size_t(0)
        _native.check(self.lib.drag_topk_batch_workspace_bytes(self.device, nq, k, self.dim, C.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _native.check(
                self.lib.drag_topk_batch(
                    self.device, _ptr(self.matrix), self.dtype_code, _ptr(state["shadow"]), self.n_rows, self.dim,
                    _ptr(self.row_sq), _ptr(state["inv"]), state["max_norm"], _ptr(d_queries), nq, k, metric_code,
                    self.row_id_base, _ptr(dist), _ptr(rows), _ptr(count), _ptr(status), _ptr(ws), need.value, stream,
                )
            )
        # queries the certificate did not cover (candidate overflow, NaN distances) go through the scan
        redo = torch.nonzero(status).flatten()
        self.last_batch_fallbacks = int(redo.numel())
        if redo.numel():
            d2, r2, c2 = self._topk_scan_device(d_queries[redo].contiguous(), k, metric_code)
            dist[redo], rows[redo], count[redo] = d2, r2, c2
        return dist, rows, count

    # ------------------------------------------------------------------ search
    def topk_device(self, d_queries, k: int, metric, allow_batch: bool = True):
        """Device in / device out: ``d_queries`` f64 ``[Q, dim]`` cuda tensor.

        Returns cuda tensors ``(dist f64[Q,k], rows i64[Q,k], count i32[Q])``.  Small batches run
        the float64 scan (asynchronous on the current stream); batches of ``batch_min_queries`` or
        more run the tensor-core candidate pass + float64 re-rank (same results; synchronises once
        to read the per-query status).
        """
        if k < 1:
            raise DragError(3, f"k={k} must be positive")
        code = _metric_code(metric)
        if k > MAX_K:
            # beyond the selection kernels' list size (the reference's argsort()[:limit] takes any limit,
            # embeddings_index.py:57-59): every distance from drag_distances, then a stable device sort -- same order
            # (ascending, NaN last, ties by row id), one full pass and a sort per query instead of a fused scan
            return self._topk_sort_device(d_queries, k, code)
        if allow_batch and self._use_batch(int(d_queries.shape[0]), k, code):
            return self._topk_batch_device(d_queries, k, code)
        return self._topk_scan_device(d_queries, k, code)

    def _topk_sort_device(self, d_queries, k: int, metric_code: int):
        torch = _torch()
        dev = self.matrix.device
        nq = int(d_queries.shape[0])
        k = min(k, self.n_rows)
        dist = torch.empty((nq, k), dtype=torch.float64, device=dev)
        rows = torch.empty((nq, k), dtype=torch.int64, device=dev)
        count = torch.full((nq,), k, dtype=torch.int32, device=dev)
        full = torch.empty(self.n_rows, dtype=torch.float64, device=dev)
        scratch = torch.empty(2, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        for i in range(nq):
            with torch.cuda.device(dev):
                _native.check(
                    self.lib.drag_distances(self.device, _ptr(self.matrix), self.dtype_code, self.n_rows, self.dim,
                                            _ptr(self.row_sq), _ptr(d_queries[i]), metric_code, _ptr(full), _ptr(scratch), stream)
                )
            vals, idx = torch.sort(full, stable=True)   # NaN sorts last, like np.argsort
            dist[i], rows[i] = vals[:k], idx[:k] + self.row_id_base
        return dist, rows, count

    def _topk_scan_device(self, d_queries, k: int, metric_code: int):
        torch = _torch()
        dev = self.matrix.device
        nq = int(d_queries.shape[0])
        dist = torch.empty((nq, k), dtype=torch.float64, device=dev)
        rows = torch.empty((nq, k), dtype=torch.int64, device=dev)
        count = torch.empty((nq,), dtype=torch.int32, device=dev)
        ws, ws_bytes = self._workspace(torch, dev, nq, k)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _native.check(
                self.lib.drag_topk(
                    self.device, _ptr(self.matrix), self.dtype_code, self.n_rows, self.dim,
                    _ptr(self.row_sq), _ptr(d_queries), nq, k, metric_code, self.row_id_base,
                    _ptr(dist), _ptr(rows), _ptr(count), _ptr(ws), ws_bytes, stream,
                )
            )
        return dist, rows, count

    def topk(self, queries: np.ndarray, k: int, metric) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Host in / host out: ``(dist f64[Q,k'], rows i64[Q,k'], count i32[Q])`` with
        ``k' = min(k, n_rows)``; rows are global ids, best first, ties -> lowest id."""
        torch = _torch()
        q = self._check_queries(queries)
        k_eff = min(int(k), self.n_rows)
        if k_eff <= 0 or q.shape[0] == 0:
            return (np.zeros((q.shape[0], 0)), np.zeros((q.shape[0], 0), dtype=np.int64),
                    np.zeros(q.shape[0], dtype=np.int32))
        dev = self.matrix.device
        d_q = torch.from_numpy(q).to(dev, non_blocking=True)
        dist, rows, count = self.topk_device(d_q, k_eff, metric)
        return dist.cpu().numpy(), rows.cpu().numpy(), count.cpu().numpy()

    def topk_chunks(self, queries: np.ndarray, k: int, metric):
        """Like ``topk`` but maps rows to ``(doc_id, chunk_id)`` on the device
        (embeddings_index.py:60,70-79).  Returns ``(doc i64[Q,k'], chunk i64[Q,k'], dist)``."""
        torch = _torch()
        q = self._check_queries(queries)
        k_eff = min(int(k), self.n_rows)
        if k_eff <= 0 or q.shape[0] == 0:
            z = np.zeros((q.shape[0], 0), dtype=np.int64)
            return z, z.copy(), np.zeros((q.shape[0], 0))
        dev = self.matrix.device
        d_q = torch.from_numpy(q).to(dev, non_blocking=True)
        dist, rows, _ = self.topk_device(d_q, k_eff, metric)
        docs = torch.empty_like(rows)
        chunks = torch.empty_like(rows)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _native.check(
                self.lib.drag_rows_to_chunks(_ptr(rows), rows.numel(), _ptr(self.doc_offsets), self.n_docs,
                                             _ptr(self.chunk_ids), _ptr(docs), _ptr(chunks), stream)
            )
        packed = torch.stack((docs, chunks)).cpu().numpy()
        return packed[0], packed[1], dist.cpu().numpy()

    def distances(self, query: np.ndarray, metric) -> np.ndarray:
        """``ENUM_TO_METRIC[metric](query, docs)``: float64 distance of every row."""
        torch = _torch()
        q = self._check_queries(query)
        if q.shape[0] != 1:
            raise ValueError("distances() takes exactly one query vector")
        dev = self.matrix.device
        out = torch.empty(self.n_rows, dtype=torch.float64, device=dev)
        if self.n_rows == 0:
            return out.cpu().numpy()
        scratch = torch.empty(2, dtype=torch.float64, device=dev)
        d_q = torch.from_numpy(q).to(dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _native.check(
                self.lib.drag_distances(self.device, _ptr(self.matrix), self.dtype_code, self.n_rows, self.dim,
                                        _ptr(self.row_sq), _ptr(d_q), _metric_code(metric), _ptr(out),
                                        _ptr(scratch), stream)
            )
        return out.cpu().numpy()


def merge_topk_device(lib, device: int, dist, rows, count, k: int):
    """Merge gathered per-shard results ``[S, Q, k]`` (cuda tensors) into the global top-k."""
    torch = _torch()
    dev = dist.device
    n_shards, nq = int(dist.shape[0]), int(dist.shape[1])
    out_d = torch.empty((nq, k), dtype=torch.float64, device=dev)
    out_r = torch.empty((nq, k), dtype=torch.int64, device=dev)
    out_c = torch.empty((nq,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _native.check(
            lib.drag_topk_merge(device, _ptr(dist), _ptr(rows), _ptr(count), n_shards, nq, k,
                                _ptr(out_d), _ptr(out_r), _ptr(out_c), stream)
        )
    return out_d, out_r, out_c
