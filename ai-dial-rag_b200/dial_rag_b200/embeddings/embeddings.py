"""bge-small-en embeddings on the B200 -- drop-in for aidial_rag/embeddings/embeddings.py.

Same module surface: ``bge_embedding`` (an ``AsyncEmbeddings``), ``bge_embedding_impl()``,
``build_embeddings(texts, stageio)``, ``EMBEDDING_LENGTH``, ``EMBEDDINGS_BATCH_SIZE``,
``BGE_EMBEDDINGS_MODEL_NAME_OR_PATH``, ``BGE_EMBEDDINGS_DEVICE``.

What changes is the execution: ``bge_embedding_impl()`` returns a ``B200BgeEmbeddings``
whose ``embed_documents`` / ``embed_query`` keep the text preparation of langchain's
``HuggingFaceBgeEmbeddings`` (newlines -> spaces; query instruction prefix) and
sentence-transformers' truncation, tokenise on the host and run the hand-written CUDA
encoder (packed batches, no padding) through the C ABI.

Differences from the reference, all deliberate:
  * the model is created on first use, not at import (embeddings.py:69 runs a forward at
    import to learn the width); ``EMBEDDING_LENGTH`` is the constant 384;
  * ``AsyncEmbeddings.embed_documents/embed_query`` work synchronously instead of raising
    ``NotImplementedError`` (embeddings.py:73-77) -- ``SemanticRetriever._get_relevant_documents``
    calls the sync form (semantic_retriever.py:49);
  * ``aembed_documents_numpy`` returns views of one float32 matrix, skipping the
    ``List[float]`` round trip the reference marks as TODO (embeddings.py:87).
"""

from __future__ import annotations

import asyncio
import logging
import os
import threading
from itertools import chain
from typing import Iterable, List, Mapping, Optional, Sequence

import numpy as np

from dial_rag_b200.batched import TqdmProgressBar, batched_map_with_progress, chunked  # noqa: F401 (re-export)
from dial_rag_b200.embeddings.detect_device import DeviceType, detect_device
from dial_rag_b200.embeddings.encoder import BGE_SMALL, B200Encoder, EncoderShape, load_model_dir
from dial_rag_b200.embeddings.tokenizer import WordPieceTokenizer
from dial_rag_b200.resources.cpu_pools import (
    run_in_indexing_cpu_pool,
    run_in_indexing_embeddings_pool,
    run_in_query_embeddings_pool,
)

try:  # use langchain's base class when the package lives inside a Dial RAG deployment
    from langchain.schema.embeddings import Embeddings  # type: ignore
except Exception:  # noqa: BLE001
    try:
        from langchain_core.embeddings import Embeddings  # type: ignore
    except Exception:  # noqa: BLE001

        class Embeddings:  # type: ignore[no-redef]
            """Minimal stand-in for langchain's ``Embeddings`` interface."""

            def embed_documents(self, texts: List[str]) -> List[List[float]]:
                raise NotImplementedError

            def embed_query(self, text: str) -> List[float]:
                raise NotImplementedError


logger = logging.getLogger(__name__)

# The reference uses 128 ("works faster on CPU with openvino", embeddings.py:24-26).  One outer
# batch is one executor submission and one packed GPU forward; 1024 chunks keep the B200 busy
# (>= 1k 128-row tiles per GEMM) while still interleaving requests batch by batch.
EMBEDDINGS_BATCH_SIZE = int(os.environ.get("DIAL_RAG_B200_EMBEDDINGS_BATCH_SIZE", "1024"))

BGE_EMBEDDINGS_MODEL_NAME_OR_PATH = os.environ.get("BGE_EMBEDDINGS_MODEL_PATH", "epam/bge-small-en")
BGE_EMBEDDINGS_DEVICE = os.environ.get("BGE_EMBEDDINGS_DEVICE", DeviceType.AUTO)

# langchain_community.embeddings.huggingface.DEFAULT_QUERY_BGE_INSTRUCTION_EN
DEFAULT_QUERY_BGE_INSTRUCTION_EN = "Represent this question for searching relevant passages: "

EMBEDDING_LENGTH = BGE_SMALL.hidden


class B200BgeEmbeddings(Embeddings):
    """``HuggingFaceBgeEmbeddings`` look-alike backed by the CUDA encoder."""

    query_instruction: str = DEFAULT_QUERY_BGE_INSTRUCTION_EN
    embed_instruction: str = ""

    def __init__(self, weights: Mapping[str, object], tokenizer: WordPieceTokenizer,
                 shape: EncoderShape = BGE_SMALL, device: int = 0, max_tokens: int = 262144):
        self.tokenizer = tokenizer
        self.client = B200Encoder(weights, shape=shape, device=device, max_tokens=max_tokens)

    @classmethod
    def from_model_dir(cls, path: str, device: int = 0, **kw) -> "B200BgeEmbeddings":
        return cls(load_model_dir(path), WordPieceTokenizer.from_model_dir(path), device=device, **kw)

    def tokenize_documents(self, texts: Sequence[str], n_threads: int = 0):
        """Host half of ``embed_documents``: text preparation + WordPiece -> packed ``(ids, cu_seqlens)``."""
        return self.tokenizer.encode_packed([self.embed_instruction + t.replace("\n", " ") for t in texts], n_threads=n_threads)

    def embed_packed_numpy(self, ids: np.ndarray, cu_seqlens: np.ndarray) -> np.ndarray:
        """Device half: packed ids -> float32 ``[n, 384]`` on the host."""
        if len(cu_seqlens) <= 1:
            return np.zeros((0, EMBEDDING_LENGTH), dtype=np.float32)
        return self.client.embed_packed(ids, cu_seqlens)

    def embed_documents_numpy(self, texts: Sequence[str]) -> np.ndarray:
        return self.embed_packed_numpy(*self.tokenize_documents(texts))

    def embed_documents(self, texts: List[str]) -> List[List[float]]:
        return self.embed_documents_numpy(texts).tolist()

    def embed_query(self, text: str) -> List[float]:
        prepared = self.query_instruction + text.replace("\n", " ")
        return self.client.embed_packed(*self.tokenizer.encode_packed([prepared], n_threads=1))[0].tolist()

    def embed_queries_numpy(self, texts: Sequence[str]) -> np.ndarray:
        """Many queries in ONE packed forward (the batched retriever entry points): float32 ``[n, 384]``, row i
        bit-identical to ``embed_query(texts[i])`` (an embedding does not depend on the batch composition)."""
        prepared = [self.query_instruction + t.replace("\n", " ") for t in texts]
        return self.embed_packed_numpy(*self.tokenizer.encode_packed(prepared, n_threads=1 if len(prepared) < 16 else 0))


_impl: Optional[B200BgeEmbeddings] = None
_impl_lock = threading.Lock()


def configure(impl: Optional[B200BgeEmbeddings]) -> None:
    """Install (or clear) the process-wide encoder, e.g. with seeded weights in tests/bench."""
    global _impl
    with _impl_lock:
        _impl = impl


def bge_embedding_impl() -> B200BgeEmbeddings:
    """Process-wide encoder, created on first use (reference: ``@cache``, embeddings.py:52-66)."""
    global _impl
    if _impl is None:
        with _impl_lock:
            if _impl is None:
                device = detect_device(BGE_EMBEDDINGS_DEVICE)
                if device == DeviceType.CPU:
                    raise RuntimeError(
                        "dial_rag_b200 has no CPU execution path: BGE_EMBEDDINGS_DEVICE must be auto/cuda/b200 "
                        "on a machine with an sm_100 GPU")
                if not os.path.isdir(BGE_EMBEDDINGS_MODEL_NAME_OR_PATH):
                    raise FileNotFoundError(
                        f"BGE_EMBEDDINGS_MODEL_PATH={BGE_EMBEDDINGS_MODEL_NAME_OR_PATH!r} is not a directory with "
                        "model.safetensors + tokenizer.json (no hub download: the deployment image bakes the model in, "
                        "reference Dockerfile:61)")
                logger.info("BGE embeddings device: b200 (cuda:%s)", os.environ.get("DIAL_RAG_B200_DEVICE", "0"))
                _impl = B200BgeEmbeddings.from_model_dir(
                    BGE_EMBEDDINGS_MODEL_NAME_OR_PATH, device=int(os.environ.get("DIAL_RAG_B200_DEVICE", "0")))
    return _impl


class AsyncEmbeddings(Embeddings):
    def embed_documents(self, texts: List[str]) -> List[List[float]]:
        return bge_embedding_impl().embed_documents(texts)

    def embed_query(self, text: str) -> List[float]:
        return bge_embedding_impl().embed_query(text)

    async def aembed_documents(self, texts: List[str]) -> List[List[float]]:
        return await run_in_indexing_embeddings_pool(bge_embedding_impl().embed_documents, texts)

    async def aembed_documents_numpy(self, texts: List[str]) -> List[np.ndarray]:
        matrix = await run_in_indexing_embeddings_pool(bge_embedding_impl().embed_documents_numpy, texts)
        return list(matrix)  # float32 row views, shape (384,)

    async def aembed_query(self, text: str) -> List[float]:
        return await run_in_query_embeddings_pool(bge_embedding_impl().embed_query, text)

    def embed_queries_numpy(self, texts: List[str]) -> np.ndarray:
        return bge_embedding_impl().embed_queries_numpy(texts)

    async def aembed_queries_numpy(self, texts: List[str]) -> np.ndarray:
        return await run_in_query_embeddings_pool(bge_embedding_impl().embed_queries_numpy, texts)


bge_embedding = AsyncEmbeddings()


# host threads one tokenisation call may use (several indexing requests can be in flight: the reference bounds its
# CPU work with the indexing pool, resources/cpu_pools.py:37-59)
TOKENIZER_THREADS = int(os.environ.get("DIAL_RAG_B200_TOKENIZER_THREADS", str(max(1, min(8, (os.cpu_count() or 2) // 2)))))


async def build_embeddings(texts: Iterable[str], stageio):
    """Embed ``texts`` in order, batch by batch, with progress lines (embeddings.py:102-108).

    Same contract as the reference (ordered results, ONE batch at a time on the indexing-embeddings worker,
    tqdm keep-alive lines); the only addition is that batch i+1 is tokenised -- on the bounded indexing CPU pool, with a
    bounded thread count -- while batch i is on the GPU, so the host half never leaves the device idle."""
    impl = bge_embedding_impl()
    batches = list(chunked(texts, EMBEDDINGS_BATCH_SIZE))
    results = []

    def tokenize(batch):
        return impl.tokenize_documents(batch, n_threads=TOKENIZER_THREADS)

    ahead = asyncio.ensure_future(run_in_indexing_cpu_pool(tokenize, batches[0])) if batches else None
    try:
        for i, _ in enumerate(TqdmProgressBar(iterable=batches, file=stageio)):
            ids, cu = await ahead
            ahead = asyncio.ensure_future(run_in_indexing_cpu_pool(tokenize, batches[i + 1])) if i + 1 < len(batches) else None
            matrix = await run_in_indexing_embeddings_pool(impl.embed_packed_numpy, ids, cu)   # strictly one batch in flight
            results.append(list(matrix))  # float32 row views, shape (384,)
    finally:
        if ahead is not None and not ahead.done():
            ahead.cancel()
        if ahead is not None:
            try:
                await ahead        # retrieve the outcome: no "exception was never retrieved", no stray work
            except BaseException:  # noqa: BLE001 - cancelled, or failed after the batch that raised
                pass
    return chain.from_iterable(results)
