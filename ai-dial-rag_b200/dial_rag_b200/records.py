"""Record / container types at the boundary of the hot path.

When the reference package (``aidial_rag``) and its container libraries are
importable -- i.e. this package is plugged into a Dial RAG deployment -- the
reference's own classes are used, so persisted ``DocumentRecord`` objects
(aidial_rag/document_record.py:32-52, FORMAT_VERSION 12) stay interchangeable.
Otherwise protocol-compatible stand-ins with the same attribute names are defined.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from enum import StrEnum
from typing import Any, Dict, List, TypedDict

import numpy as np

try:  # pragma: no cover - exercised only inside a Dial RAG deployment
    from aidial_rag.document_record import Chunk, ItemEmbeddings, MultiEmbeddings
    from aidial_rag.index_record import RetrievalType, to_metadata_doc
    from langchain.schema import Document

    USING_REFERENCE_TYPES = True
except Exception:  # noqa: BLE001 - any import problem means "not deployed inside Dial RAG"
    USING_REFERENCE_TYPES = False

    try:
        from langchain_core.documents import Document  # type: ignore
    except Exception:  # noqa: BLE001

        @dataclass
        class Document:  # type: ignore[no-redef]
            """langchain ``Document`` look-alike (value equality on both fields)."""

            page_content: str
            metadata: Dict[str, Any] = field(default_factory=dict)

    class RetrievalType(StrEnum):  # aidial_rag/index_record.py:18-20
        TEXT = "text"
        IMAGE = "image"

    class ChunkMetadata(TypedDict):  # aidial_rag/index_record.py:23-26
        doc_id: int
        chunk_id: int
        retrieval_type: RetrievalType

    def to_metadata_doc(doc_id: int, chunk_id: int, retrieval_type: RetrievalType) -> Document:
        """aidial_rag/index_record.py:29-38 -- EnsembleRetriever keys on page_content."""
        return Document(
            page_content=f"{doc_id}_{chunk_id}",
            metadata=ChunkMetadata(doc_id=doc_id, chunk_id=chunk_id, retrieval_type=retrieval_type),
        )

    @dataclass
    class Chunk:  # aidial_rag/document_record.py:15-24
        text: str
        metadata: dict = field(default_factory=dict)
        id: str | None = None

    @dataclass(eq=False)
    class ItemEmbeddings:  # aidial_rag/document_record.py:32-36
        """Embeddings of one item (chunk or page): float32 ``[n_i, dim]``."""

        embeddings: np.ndarray
        id: str | None = None

    class MultiEmbeddings(List[ItemEmbeddings]):  # aidial_rag/document_record.py:39
        pass
