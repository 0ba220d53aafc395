"""Generate tests/golden/index_build.npz by running the REAL reference index builders / packers.

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden_index_build.py

Executes the reference's own, unmodified ``create_index_by_page``, ``create_index_by_chunk``,
``pack_multi_embeddings`` and ``pack_simple_embeddings``
(aidial_rag/retrievers/embeddings_index.py:101-164, loaded through ``oracle/ref_shims.py``; the
``Chunk`` / ``ItemEmbeddings`` containers come from the reference's ``document_record.py`` over the same
docarray stand-ins) on seeded inputs and stores inputs + outputs.  ``tests/test_index_builders.py`` runs this
package's builders on the same inputs and compares ids, values, shapes and dtypes exactly.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_shims import load_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# (name, rows per item, dim, dtype of the item arrays)
CHUNK_CASES = [
    ("simple", [1] * 9, 16, "float32"),
    ("multi_row", [1, 3, 0, 2, 1, 0, 0, 5, 1], 24, "float32"),
    ("float64_items", [2, 1, 1], 8, "float64"),
    ("all_empty", [0, 0, 0], 8, "float32"),
    ("single", [4], 384, "float32"),
]
# (name, rows per page, page number (1-based) of every chunk, dim)
PAGE_CASES = [
    ("pages", [2, 0, 3, 1, 1, 4, 2], [1, 1, 2, 3, 3, 3, 4, 6, 6, 7, 5, 1], 16),
    ("one_page", [3], [1, 1, 1, 1], 32),
    ("empty_pages_only", [0, 0], [1, 2, 2], 8),
]
# (name, page index of every embedding, number of pages, dim)
PACK_MULTI_CASES = [
    ("grouped", [0, 0, 2, 5, 5, 5, 1, 0], 7, 16),
    ("none", [], 3, 8),
]


def item_arrays(rng, rows_per_item, dim, dtype):
    return [rng.standard_normal((n, dim)).astype(dtype) for n in rows_per_item]


def main() -> None:
    _, index_mod, _ = load_reference()
    import aidial_rag.document_record as rec  # the reference's containers (over the docarray stand-ins)

    out = {}
    rng = np.random.default_rng(2024)

    for name, rows, dim, dtype in CHUNK_CASES:
        arrays = item_arrays(rng, rows, dim, dtype)
        multi = rec.MultiEmbeddings([rec.ItemEmbeddings(embeddings=a) for a in arrays])
        got = index_mod.create_index_by_chunk(multi)
        for i, a in enumerate(arrays):
            out[f"chunk:{name}:in{i}"] = a
        out[f"chunk:{name}:n"] = np.array(len(arrays))
        out[f"chunk:{name}:chunk_ids"] = got.chunk_ids
        out[f"chunk:{name}:embeddings"] = got.embeddings
    none = index_mod.create_index_by_chunk(None)
    out["chunk:none:chunk_ids"], out["chunk:none:embeddings"] = none.chunk_ids, none.embeddings

    for name, rows, pages, dim in PAGE_CASES:
        arrays = item_arrays(rng, rows, dim, "float32")
        multi = rec.MultiEmbeddings([rec.ItemEmbeddings(embeddings=a) for a in arrays])
        chunks = [rec.Chunk(text=f"c{i}", metadata={"page_number": p}) for i, p in enumerate(pages)]
        got = index_mod.create_index_by_page(chunks, multi)
        for i, a in enumerate(arrays):
            out[f"page:{name}:in{i}"] = a
        out[f"page:{name}:n"] = np.array(len(arrays))
        out[f"page:{name}:page_numbers"] = np.array(pages)
        out[f"page:{name}:chunk_ids"] = got.chunk_ids
        out[f"page:{name}:embeddings"] = got.embeddings
    none = index_mod.create_index_by_page([], None)
    out["page:none:chunk_ids"], out["page:none:embeddings"] = none.chunk_ids, none.embeddings

    for name, indexes, n_pages, dim in PACK_MULTI_CASES:
        embs = [rng.standard_normal(dim).astype(np.float32) for _ in indexes]
        multi = index_mod.pack_multi_embeddings(indexes, embs, n_pages)
        out[f"packmulti:{name}:indexes"] = np.array(indexes, dtype=np.int64)
        out[f"packmulti:{name}:in"] = np.array(embs, dtype=np.float32).reshape(len(embs), dim)
        out[f"packmulti:{name}:pages"] = np.array(n_pages)
        assert len(multi) == n_pages
        for p, item in enumerate(multi):
            out[f"packmulti:{name}:out{p}"] = np.asarray(item.embeddings)

    embs = [rng.standard_normal(12) for _ in range(5)]   # float64 in: the packer casts to float32
    multi = index_mod.pack_simple_embeddings(embs)
    out["packsimple:in"] = np.array(embs)
    for i, item in enumerate(multi):
        out[f"packsimple:out{i}"] = np.asarray(item.embeddings)
    # ... and through create_index_by_chunk, as SemanticRetriever.from_doc_records does (semantic_retriever.py:30-34)
    flat = index_mod.create_index_by_chunk(multi)
    out["packsimple:chunk_ids"], out["packsimple:embeddings"] = flat.chunk_ids, flat.embeddings

    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, "index_build.npz"), **out)
    print("wrote index_build.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
