"""CPU ORACLE support (test infrastructure): run the REAL reference search code.

The reference's ``aidial_rag/retrievers/embeddings_metrics.py`` imports only
numpy + torch, and ``embeddings_index.py`` additionally needs ``docarray`` and
``langchain`` purely as *containers* (``DocList``, ``NdArray``, ``Document``).
Neither package is installed in the authoring container, so this module
registers minimal stand-ins for those container types in ``sys.modules`` and
then imports the reference's own, unmodified source files from
``/root/reference``.  All arithmetic, sorting and tie-breaking executed through
``load_reference()`` is therefore the reference's, not ours.

Used only by ``oracle/make_golden_search.py`` (authoring container; the GPU box
has no ``/root/reference``) to produce ``tests/golden/search_*.json``.
"""

from __future__ import annotations

import os
import sys
import types
from dataclasses import dataclass, field
from typing import Any

REFERENCE_ROOT = os.environ.get("DIAL_RAG_REFERENCE_ROOT", "/root/reference")


@dataclass
class _Document:  # stand-in for langchain.schema.Document (value semantics)
    page_content: str
    metadata: dict = field(default_factory=dict)


class _BaseDoc:  # stand-in for docarray.BaseDoc: keyword constructor only
    def __init__(self, **kw: Any) -> None:
        for k, v in kw.items():
            setattr(self, k, v)


class _DocList(list):  # stand-in for docarray.DocList[T]
    def __class_getitem__(cls, item):
        return cls


class _NdArray:  # stand-in for docarray.typing.NdArray
    def __class_getitem__(cls, item):
        return cls

    def __new__(cls, shape=None, buffer=None, dtype=None):
        return buffer


def _install_stubs() -> None:
    if "docarray" not in sys.modules:
        docarray = types.ModuleType("docarray")
        docarray.BaseDoc = _BaseDoc
        docarray.DocList = _DocList
        typing_mod = types.ModuleType("docarray.typing")
        typing_mod.NdArray = _NdArray
        typing_mod.ID = str
        docarray.typing = typing_mod
        sys.modules["docarray"] = docarray
        sys.modules["docarray.typing"] = typing_mod
    if "langchain" not in sys.modules:
        langchain = types.ModuleType("langchain")
        schema = types.ModuleType("langchain.schema")
        schema.Document = _Document
        langchain.schema = schema
        sys.modules["langchain"] = langchain
        sys.modules["langchain.schema"] = schema


def reference_available() -> bool:
    return os.path.isfile(
        os.path.join(
            REFERENCE_ROOT, "aidial_rag", "retrievers", "embeddings_index.py"
        )
    )


def load_reference():
    """Returns (embeddings_metrics, embeddings_index, index_record) modules."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import aidial_rag.index_record as index_record
    import aidial_rag.retrievers.embeddings_index as embeddings_index
    import aidial_rag.retrievers.embeddings_metrics as embeddings_metrics

    return embeddings_metrics, embeddings_index, index_record
