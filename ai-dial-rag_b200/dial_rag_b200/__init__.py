"""dial_rag_b200 -- B200-native drop-in for Dial RAG's semantic-retriever hot path.

Mirrors the reference's module layout for the path it replaces
(``aidial_rag.embeddings.embeddings``, ``aidial_rag.retrievers.embeddings_index``,
``aidial_rag.retrievers.embeddings_metrics``, ``aidial_rag.retrievers.semantic_retriever``,
``aidial_rag.batched``, ``aidial_rag.resources.cpu_pools``).  All arithmetic runs in
``libdrag_b200.so`` (hand-written CUDA for sm_100a) behind the C ABI in
``include/drag_b200.h``; there is no CPU fallback.
"""

__version__ = "0.1.0"
