"""Semantic retriever -- drop-in for aidial_rag/retrievers/semantic_retriever.py.

``from_doc_records`` flattens each document's ``embeddings_index`` (one row per chunk)
into a ``DocIndex`` and builds an ``EmbeddingsIndex`` with the DEFAULT metric (squared
euclidean, semantic_retriever.py:36-40); a query is embedded with the query instruction,
turned into a float64 vector (``np.array(list_of_python_floats)``, :49/:53) and searched;
``build_index`` = ``build_embeddings`` + ``pack_simple_embeddings`` (:58-66).
"""

from __future__ import annotations

import asyncio
import logging
import sys
from time import perf_counter
from typing import Any, List

import numpy as np

from dial_rag_b200.embeddings import embeddings as _emb
from dial_rag_b200.records import Document, MultiEmbeddings, RetrievalType
from dial_rag_b200.retrievers.embeddings_index import (
    EmbeddingsIndex,
    pack_simple_embeddings,
)

try:  # pragma: no cover - only inside a Dial RAG deployment
    from langchain.schema import BaseRetriever  # type: ignore

    _HAVE_LANGCHAIN = True
except Exception:  # noqa: BLE001
    _HAVE_LANGCHAIN = False

    class BaseRetriever:  # type: ignore[no-redef]
        """The two entry points langchain's ``BaseRetriever`` gives callers."""

        def __init__(self, **fields: Any):
            for k, v in fields.items():
                setattr(self, k, v)

        def invoke(self, query: str, *args, **kwargs) -> List[Document]:
            return self._get_relevant_documents(query)

        async def ainvoke(self, query: str, *args, **kwargs) -> List[Document]:
            return await self._aget_relevant_documents(query)

        def batch(self, inputs: List[str], *args, **kwargs) -> List[List[Document]]:
            return [self.invoke(q) for q in inputs]

        async def abatch(self, inputs: List[str], *args, **kwargs) -> List[List[Document]]:
            return list(await asyncio.gather(*(self.ainvoke(q) for q in inputs)))


logger = logging.getLogger(__name__)


class SemanticRetriever(BaseRetriever):
    index: Any  # EmbeddingsIndex (Any: pydantic-based BaseRetriever must not validate it)

    if _HAVE_LANGCHAIN:  # pragma: no cover
        model_config = {"arbitrary_types_allowed": True}

    @classmethod
    def from_doc_records(cls, document_records: List[Any], k: int = 1) -> "SemanticRetriever":
        # same documents as semantic_retriever.py:30-34; the flattened matrix stays in HBM between requests
        # (keyed by the identity of the persisted per-document embeddings, see ResidentIndexCache)
        sources = [doc.embeddings_index for doc in document_records if doc.embeddings_index]
        return cls(index=EmbeddingsIndex.from_sources(RetrievalType.TEXT, sources, limit=k))

    def _find_relevant_documents(self, query_emb: np.ndarray) -> List[Document]:
        return self.index.find(query=query_emb)

    def _get_relevant_documents(self, query: str, *args, **kwargs) -> List[Document]:
        query_emb = np.array(_emb.bge_embedding.embed_query(query))  # float64, as in the reference
        return self._find_relevant_documents(query_emb)

    async def _aget_relevant_documents(self, query: str, *args, **kwargs) -> List[Document]:
        query_emb = np.array(await _emb.bge_embedding.aembed_query(query))
        return await asyncio.get_running_loop().run_in_executor(None, self._find_relevant_documents, query_emb)

    # langchain's Runnable.batch / abatch fan a list of queries out to N single invocations (N forwards at batch 1 and
    # N passes over the matrix).  Here: ONE packed forward for all queries, ONE find_batch (eval/eval_retriever.py:97 is
    # the reference's caller).  Per-call callbacks/config of the Runnable protocol are not used on this path.
    def batch(self, inputs: List[str], config: Any = None, **kwargs: Any) -> List[List[Document]]:
        inputs = list(inputs)
        if not inputs:
            return []
        queries = _emb.bge_embedding.embed_queries_numpy(inputs).astype(np.float64)
        return self.index.find_batch(queries)

    async def abatch(self, inputs: List[str], config: Any = None, **kwargs: Any) -> List[List[Document]]:
        inputs = list(inputs)
        if not inputs:
            return []
        queries = (await _emb.bge_embedding.aembed_queries_numpy(inputs)).astype(np.float64)
        return await asyncio.get_running_loop().run_in_executor(None, self.index.find_batch, queries)

    @staticmethod
    async def build_index(chunks: List[Any], stageio=sys.stderr) -> MultiEmbeddings:
        stageio.write("Building Semantic indexes started\n")  # utils.timed_block, utils.py:26-34
        start = perf_counter()
        try:
            logger.debug("Building Semantic indexes.")
            embeddings = await _emb.build_embeddings((chunk.text for chunk in chunks), stageio)
            return pack_simple_embeddings(embeddings)
        finally:
            stageio.write(f"Building Semantic indexes took {perf_counter() - start:.2f}s\n")
