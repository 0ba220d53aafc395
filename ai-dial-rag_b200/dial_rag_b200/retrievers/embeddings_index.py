"""Embeddings index with GPU search -- drop-in for aidial_rag/retrievers/embeddings_index.py.

Same public surface (``DocIndex``, ``EmbeddingsIndex(retrieval_type, indexes, metric,
limit).find/find_in_doc``, ``create_index_by_chunk/page``, ``pack_simple/multi_embeddings``,
``to_ndarray``) and the same results; the execution differs:

* the per-document matrices are concatenated once, in document order, into ONE
  row-major device matrix with a document-offset table.  The reference's
  "per-document stable top-k, then stable top-k of the concatenated winners"
  (embeddings_index.py:62-89) equals one global stable top-k over that concatenation
  (ties -> earlier document, then lower row == lowest global row id), so a single
  fused scan answers ``find``;
* scoring + selection is ``drag_topk`` (float64 accumulation like numpy's path,
  NaN last like ``np.argsort``), row -> (doc_id, chunk_id) is ``drag_rows_to_chunks``.
"""

from __future__ import annotations

import os
import threading
import uuid
import weakref
from collections import OrderedDict
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import numpy.typing as npt

from dial_rag_b200.device_index import DeviceMatrix
from dial_rag_b200.records import (
    Chunk,
    Document,
    ItemEmbeddings,
    MultiEmbeddings,
    RetrievalType,
    to_metadata_doc,
)
from dial_rag_b200.retrievers.embeddings_metrics import Metric

__all__ = [
    "DocIndex", "EmbeddingsIndex", "Metric", "create_index_by_chunk", "create_index_by_page",
    "pack_multi_embeddings", "pack_simple_embeddings", "to_ndarray",
]


class DocIndex:
    """Row -> chunk map and embedding rows of one document (embeddings_index.py:14-30)."""

    chunk_ids: npt.NDArray[np.int64]
    embeddings: np.ndarray

    def __init__(self, chunk_ids: npt.NDArray[np.int64] | None = None, embeddings: np.ndarray | None = None):
        self.chunk_ids = np.array([], dtype=np.int64) if chunk_ids is None else chunk_ids
        self.embeddings = np.array([], dtype=np.float32) if embeddings is None else embeddings
        self._device: Optional[DeviceMatrix] = None

    def __len__(self) -> int:
        return len(self.embeddings)


def _stack_documents(doc_indexes: Sequence[DocIndex]):
    """Concatenate non-empty documents in order; offsets repeat for empty ones."""
    mats, ids, offsets = [], [], [0]
    dim = None
    for doc in doc_indexes:
        n = len(doc.embeddings)
        if n:
            emb = np.asarray(doc.embeddings)
            if emb.ndim != 2:
                raise ValueError(f"document embeddings must be 2-d, got shape {emb.shape}")
            if dim is None:
                dim = emb.shape[1]
            elif emb.shape[1] != dim:
                raise ValueError(f"embedding dimension mismatch between documents: {dim} vs {emb.shape[1]}")
            if len(doc.chunk_ids) != n:
                raise ValueError("chunk_ids and embeddings must have the same length")
            mats.append(np.ascontiguousarray(emb, dtype=np.float32))
            ids.append(np.asarray(doc.chunk_ids, dtype=np.int64))
        offsets.append(offsets[-1] + n)
    if not mats:
        return None, None, offsets
    return np.concatenate(mats), np.concatenate(ids), offsets



UID_PREFIX = "drag-"   # ItemEmbeddings.id of the first item of an index built by this package: "drag-<uuid4 hex>"


def _tag(a: np.ndarray) -> int:
    return hash(np.ascontiguousarray(a).tobytes())


def source_key(src) -> tuple:
    """A key of one document's persisted ``MultiEmbeddings`` that is STABLE ACROSS DESERIALISATION.

    The reference's record cache holds the *serialised bytes* (index_storage.py:56-66, ``LRUCacheStorage``) and
    ``DocumentRecord.from_bytes`` runs on every load (index_storage.py:136), so every request sees fresh Python
    objects: object identity cannot key anything.  What does survive pickle + gzip is the content:
      * indexes built by this package carry a build id in a field the reference persists but never reads
        (``ItemEmbeddings.id`` of the first item, document_record.py:32-36): the key is that id plus the item
        count, the width and a fingerprint of the first and last item -- O(1) per request;
      * anything else (indexes written by the stock reference) is keyed by a blake2b digest of all rows.
    """
    n = len(src)
    if n == 0:
        return ("empty",)
    first = np.asarray(src[0].embeddings)
    uid = getattr(src[0], "id", None)
    if isinstance(uid, str) and uid.startswith(UID_PREFIX):
        last = np.asarray(src[-1].embeddings)
        return ("uid", uid, n, first.shape, last.shape, str(first.dtype), _tag(first), _tag(last))
    import hashlib

    h = hashlib.blake2b(digest_size=16)
    rows = 0
    for item in src:
        a = np.ascontiguousarray(item.embeddings)
        rows += len(a)
        h.update(a.data if a.size else b"")
        h.update(b"|")
    return ("content", h.digest(), n, rows, str(first.dtype))


class ResidentIndexCache:
    """Device-resident indexes kept ACROSS requests.

    The reference rebuilds its flat host arrays on every request (retrieval_chain.py:264-271 ->
    ``from_doc_records``); with the matrix in HBM that would mean re-flattening and re-uploading it per request.
    Entries are keyed by VALUE (``source_key``: build id or content digest of every document, in order), so a
    request whose records were freshly deserialised from the same stored bytes hits.  Least-recently-used entries
    are evicted beyond ``max_bytes`` (``DRAG_INDEX_CACHE_MB``, default 16384; the bytes an entry pins are asked of
    the entry at eviction time, so a scoring copy made after insertion counts) or ``max_entries``.  No weak
    references, no callbacks: nothing can re-enter the lock.
    """

    def __init__(self, max_entries: int = 16, max_bytes: Optional[int] = None):
        if max_bytes is None:
            max_bytes = int(os.environ.get("DRAG_INDEX_CACHE_MB", "16384")) << 20
        self.max_entries, self.max_bytes = int(max_entries), int(max_bytes)
        self._entries: "OrderedDict[tuple, tuple]" = OrderedDict()   # key -> (value, nbytes callable)
        self._lock = threading.RLock()
        self.hits = self.misses = 0

    @staticmethod
    def _key(sources: Sequence[object], extra: tuple) -> tuple:
        return tuple(source_key(s) for s in sources) + ("|",) + tuple(extra)

    def get(self, sources: Sequence[object], extra: tuple = (), key: Optional[tuple] = None):
        key = self._key(sources, extra) if key is None else key
        with self._lock:
            entry = self._entries.get(key)
            if entry is not None:
                self._entries.move_to_end(key)
                self.hits += 1
                return entry[0]
            self.misses += 1
            return None

    def put(self, sources: Sequence[object], extra: tuple, value, nbytes, key: Optional[tuple] = None) -> bool:
        """``nbytes``: an int, or a callable returning the bytes the entry pins right now."""
        key = self._key(sources, extra) if key is None else key
        size = nbytes if callable(nbytes) else (lambda n=int(nbytes): n)
        with self._lock:
            self._entries[key] = (value, size)
            self._entries.move_to_end(key)
            while len(self._entries) > self.max_entries or (
                len(self._entries) > 1 and sum(e[1]() for e in self._entries.values()) > self.max_bytes
            ):
                self._entries.popitem(last=False)
        return True

    def nbytes(self) -> int:
        with self._lock:
            return sum(e[1]() for e in self._entries.values())

    def clear(self) -> None:
        with self._lock:
            self._entries.clear()

    def __len__(self) -> int:
        return len(self._entries)


RESIDENT_INDEXES = ResidentIndexCache()


class EmbeddingsIndex:
    retrieval_type: RetrievalType
    doc_indexes: List[DocIndex]
    metric: str
    limit: int

    def __init__(
        self,
        retrieval_type: RetrievalType,
        indexes: List[DocIndex],
        metric: Metric = Metric.SQEUCLIDEAN_DIST,
        limit: int = 1,
        device: Optional[int] = None,
        storage: str = "f32",
    ):
        self.retrieval_type = retrieval_type
        self.metric = metric
        self.limit = limit
        self._doc_indexes = indexes
        self._sources: Optional[Sequence[object]] = None
        self._device_id = device
        self._storage = storage
        self._resident: Optional[DeviceMatrix] = None
        self._resident_empty = False
        self._lock = threading.Lock()

    @property
    def doc_indexes(self) -> List[DocIndex]:
        """Per-document host arrays (the reference's ``indexes``); flattened lazily when the index came out of
        the resident cache -- ``find`` does not need them."""
        if self._doc_indexes is None:
            self._doc_indexes = [create_index_by_chunk(src) for src in self._sources or ()]
        return self._doc_indexes

    @classmethod
    def from_sources(
        cls,
        retrieval_type: RetrievalType,
        sources: Sequence[object],
        metric: Metric = Metric.SQEUCLIDEAN_DIST,
        limit: int = 1,
        device: Optional[int] = None,
        storage: str = "f32",
        cache: Optional["ResidentIndexCache"] = None,
    ) -> "EmbeddingsIndex":
        """An index over per-document ``MultiEmbeddings`` (one row per chunk, ``create_index_by_chunk``) whose
        device-resident matrix is shared between requests through ``cache`` (default: the process-wide one),
        keyed by the documents' build ids / content (``source_key``): records deserialised afresh from the same
        stored bytes reuse the matrix already in HBM."""
        cache = RESIDENT_INDEXES if cache is None else cache
        sources = list(sources)
        index = cls(retrieval_type, None, metric=metric, limit=limit, device=device, storage=storage)  # type: ignore[arg-type]
        index._sources = sources
        key = cache._key(sources, (storage, device))
        hit = cache.get(sources, key=key)
        if hit is not None:
            index._resident, index._resident_empty = hit
            return index
        matrix = index._matrix()   # flattens (doc_indexes) and uploads
        cache.put(sources, (storage, device), (matrix, matrix is None), (lambda: 0) if matrix is None else matrix.nbytes, key=key)
        return index

    # The matrix is uploaded on first use and then stays in HBM for the life of the
    # index object (the reference rebuilds host arrays per request, retrieval_chain.py:264-271).
    def _matrix(self) -> Optional[DeviceMatrix]:
        if self._resident is None and not self._resident_empty:
            with self._lock:
                if self._resident is None and not self._resident_empty:
                    flat, ids, offsets = _stack_documents(self.doc_indexes)
                    if flat is None:
                        self._resident_empty = True
                    else:
                        self._resident = DeviceMatrix(
                            flat, device=self._device_id, storage=self._storage,
                            chunk_ids=ids, doc_offsets=offsets,
                        )
        return self._resident

    def find_in_doc(
        self, query: np.ndarray, doc_index: DocIndex
    ) -> Tuple[npt.NDArray[np.int64], npt.NDArray[np.float64]]:
        """Top ``limit`` rows of one document: ``(chunk_ids, distances)`` (embeddings_index.py:51-60)."""
        if len(doc_index.embeddings) == 0:
            return np.array([], dtype=np.int64), np.array([], dtype=np.float64)
        if doc_index._device is None:
            doc_index._device = DeviceMatrix(
                np.asarray(doc_index.embeddings), device=self._device_id, storage=self._storage,
                chunk_ids=np.asarray(doc_index.chunk_ids, dtype=np.int64),
            )
        _, chunks, dist = doc_index._device.topk_chunks(np.asarray(query), self.limit, Metric(self.metric))
        return chunks[0], dist[0]

    def find(self, query: np.ndarray) -> List[Document]:
        """embeddings_index.py:62-89 for one query vector."""
        return self.find_batch(np.asarray(query)[None, :])[0]

    def find_batch(self, queries: np.ndarray) -> List[List[Document]]:
        """``find`` for many query vectors sharing one pass over the matrix."""
        queries = np.asarray(queries)
        matrix = self._matrix()
        if matrix is None:
            return [[] for _ in range(len(queries))]
        docs, chunks, _ = matrix.topk_chunks(queries, self.limit, Metric(self.metric))
        return [
            [
                to_metadata_doc(int(d), int(c), retrieval_type=self.retrieval_type)
                for d, c in zip(docs[i], chunks[i], strict=True)
            ]
            for i in range(len(queries))
        ]


def _get_page_index(chunk: Chunk) -> int:
    # page numbers are 1-based in chunk metadata (embeddings_index.py:92-94)
    return chunk.metadata["page_number"] - 1


try:  # resolved once: a failing import per call cost 50 us x 1M chunks in pack_embedding_matrix (bench extra.index_build)
    from docarray.typing import NdArray as _NdArray  # type: ignore
except Exception:  # noqa: BLE001
    _NdArray = None


def to_ndarray(arr: np.ndarray):
    """Wrap as docarray ``NdArray`` when docarray is present (embeddings_index.py:97-98)."""
    if _NdArray is None:
        return arr
    return _NdArray(shape=arr.shape, buffer=arr, dtype=arr.dtype)


def _rows_of(item) -> np.ndarray:
    return np.asarray(item.embeddings)


def _new_uid() -> str:
    return UID_PREFIX + uuid.uuid4().hex


# Containers built by ``pack_embedding_matrix`` keep their rows in ONE contiguous matrix (every item is a [1, dim]
# view of it).  The matrix is remembered per container (weakly, by identity: this is an in-process shortcut for the
# build -> search flow of one request, not a cache key) so that flattening is O(1) instead of a loop over n items.
_CONTIGUOUS: dict = {}


def _remember_contiguous(multi, matrix: np.ndarray) -> None:
    key = id(multi)
    try:
        _CONTIGUOUS[key] = (weakref.ref(multi, lambda _r, k=key: _CONTIGUOUS.pop(k, None)), matrix)
    except TypeError:   # container type without weak-reference support
        pass


def _contiguous_matrix(multi) -> Optional[np.ndarray]:
    """The [n, dim] matrix behind ``multi`` if it still is n ordered [1, dim] views of it (length and the first,
    middle and last item are verified: same memory, same shape)."""
    entry = _CONTIGUOUS.get(id(multi))
    if entry is None or entry[0]() is not multi:
        return None
    matrix = entry[1]
    n = len(multi)
    if n != len(matrix) or n == 0:
        return None
    for i in {0, n // 2, n - 1}:
        rows = _rows_of(multi[i])
        if rows.shape != (1, matrix.shape[1]) or rows.dtype != matrix.dtype or \
                rows.__array_interface__["data"][0] != matrix[i : i + 1].__array_interface__["data"][0]:
            return None
    return matrix


def _flatten(items: Sequence, owner_of_item, dtype) -> DocIndex:
    """Rows of ``items`` back to back + the owner id of every row; ``dtype=None`` keeps the items' dtype
    (``np.array(list_of_rows)``, embeddings_index.py:133-136), else casts (``:115-118``)."""
    lens = np.fromiter((len(_rows_of(it)) for it in items), dtype=np.int64, count=len(items))
    owners = np.repeat(np.asarray(owner_of_item, dtype=np.int64), lens)
    blocks = [_rows_of(it) for it, n in zip(items, lens) if n]
    if not blocks:
        # the reference builds np.array([]) here: float64 without a dtype, float32 with one
        return DocIndex(chunk_ids=owners, embeddings=np.array([], dtype=dtype))
    blocks = [b.reshape(len(b), -1) for b in blocks]
    flat = np.concatenate(blocks)
    return DocIndex(chunk_ids=owners, embeddings=flat if dtype is None else flat.astype(dtype, copy=False))


def create_index_by_page(chunks: Sequence[Chunk], pages_embeddings: MultiEmbeddings | None) -> DocIndex:
    """Every chunk gets all rows of its page, in chunk order (embeddings_index.py:101-118)."""
    if pages_embeddings is None:
        return DocIndex()
    pages = [pages_embeddings[_get_page_index(chunk)] for chunk in chunks]
    return _flatten(pages, np.arange(len(pages)), np.float32)


def create_index_by_chunk(chunks_embeddings: MultiEmbeddings | None) -> DocIndex:
    """Item i owns ``len(item.embeddings)`` consecutive rows (embeddings_index.py:121-136)."""
    if chunks_embeddings is None:
        return DocIndex()
    matrix = _contiguous_matrix(chunks_embeddings)
    if matrix is not None:   # n ordered [1, dim] views of one matrix: nothing to copy
        return DocIndex(chunk_ids=np.arange(len(matrix), dtype=np.int64), embeddings=matrix)
    return _flatten(chunks_embeddings, np.arange(len(chunks_embeddings)), None)


def pack_multi_embeddings(indexes: List[int], embeddings: Iterable[np.ndarray], number_of_pages: int) -> MultiEmbeddings:
    """Group embeddings by page index (embeddings_index.py:139-153)."""
    per_page: List[List[np.ndarray]] = [[] for _ in range(number_of_pages)]
    for page, emb in zip(indexes, embeddings, strict=True):
        per_page[page].append(emb)
    items = [ItemEmbeddings(embeddings=to_ndarray(np.array(rows, dtype=np.float32))) for rows in per_page]
    if items:
        items[0].id = _new_uid()
    return MultiEmbeddings(items)


def pack_simple_embeddings(embeddings: Iterable[np.ndarray]) -> MultiEmbeddings:
    """One ``[1, dim]`` float32 array per chunk (embeddings_index.py:156-164).  The first item carries a build id
    in its (persisted, otherwise unused) ``id`` field: the key of the device-resident index cache."""
    items = [ItemEmbeddings(embeddings=to_ndarray(np.array([e], dtype=np.float32))) for e in embeddings]
    if items:
        items[0].id = _new_uid()
    return MultiEmbeddings(items)


def pack_embedding_matrix(matrix: np.ndarray) -> MultiEmbeddings:
    """Zero-copy variant of ``pack_simple_embeddings`` for an ``[n, dim]`` float32 matrix: every item is a
    ``[1, dim]`` view into ``matrix``; ``create_index_by_chunk`` of the result is O(1)."""
    matrix = np.ascontiguousarray(matrix, dtype=np.float32)
    items = [ItemEmbeddings(embeddings=to_ndarray(matrix[i : i + 1])) for i in range(len(matrix))]
    if items:
        items[0].id = _new_uid()
    multi = MultiEmbeddings(items)
    _remember_contiguous(multi, matrix)
    return multi
