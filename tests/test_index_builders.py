"""``create_index_by_page`` / ``create_index_by_chunk`` / ``pack_multi_embeddings`` / ``pack_simple_embeddings``
against the outputs of the reference's own functions (aidial_rag/retrievers/embeddings_index.py:101-164, run through
oracle/ref_shims.py by oracle/make_golden_index_build.py -> tests/golden/index_build.npz): ids, values, shapes and
dtypes must be identical."""

import os

import numpy as np
import pytest

from dial_rag_b200.records import Chunk, ItemEmbeddings, MultiEmbeddings
from dial_rag_b200.retrievers.embeddings_index import (
    create_index_by_chunk,
    create_index_by_page,
    pack_multi_embeddings,
    pack_simple_embeddings,
)

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "index_build.npz"))


def _same(got: np.ndarray, want: np.ndarray, what):
    got = np.asarray(got)
    assert got.dtype == want.dtype, (what, got.dtype, want.dtype)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert np.array_equal(got, want), what


def _multi(prefix):
    n = int(Z[f"{prefix}:n"])
    return MultiEmbeddings([ItemEmbeddings(embeddings=Z[f"{prefix}:in{i}"]) for i in range(n)])


@pytest.mark.parametrize("name", ["simple", "multi_row", "float64_items", "all_empty", "single"])
def test_create_index_by_chunk_equals_reference(name):
    got = create_index_by_chunk(_multi(f"chunk:{name}"))
    _same(got.chunk_ids, Z[f"chunk:{name}:chunk_ids"], name)
    _same(got.embeddings, Z[f"chunk:{name}:embeddings"], name)


@pytest.mark.parametrize("name", ["pages", "one_page", "empty_pages_only"])
def test_create_index_by_page_equals_reference(name):
    pages = Z[f"page:{name}:page_numbers"]
    chunks = [Chunk(text=f"c{i}", metadata={"page_number": int(p)}) for i, p in enumerate(pages)]
    got = create_index_by_page(chunks, _multi(f"page:{name}"))
    _same(got.chunk_ids, Z[f"page:{name}:chunk_ids"], name)
    _same(got.embeddings, Z[f"page:{name}:embeddings"], name)


def test_none_gives_the_reference_empty_index():
    for got, prefix in ((create_index_by_chunk(None), "chunk:none"), (create_index_by_page([], None), "page:none")):
        _same(got.chunk_ids, Z[f"{prefix}:chunk_ids"], prefix)
        _same(got.embeddings, Z[f"{prefix}:embeddings"], prefix)
        assert len(got) == 0


@pytest.mark.parametrize("name", ["grouped", "none"])
def test_pack_multi_embeddings_equals_reference(name):
    indexes = Z[f"packmulti:{name}:indexes"].tolist()
    embs = list(Z[f"packmulti:{name}:in"])
    n_pages = int(Z[f"packmulti:{name}:pages"])
    got = pack_multi_embeddings(indexes, embs, n_pages)
    assert len(got) == n_pages
    for p in range(n_pages):
        _same(got[p].embeddings, Z[f"packmulti:{name}:out{p}"], (name, p))
    with pytest.raises(ValueError):
        pack_multi_embeddings(indexes + [0], embs, n_pages)   # zip(strict=True), as in the reference


def test_pack_simple_embeddings_and_flatten_equal_reference():
    got = pack_simple_embeddings(list(Z["packsimple:in"]))
    assert len(got) == 5
    for i in range(5):
        _same(got[i].embeddings, Z[f"packsimple:out{i}"], i)
    flat = create_index_by_chunk(got)
    _same(flat.chunk_ids, Z["packsimple:chunk_ids"], "ids")
    _same(flat.embeddings, Z["packsimple:embeddings"], "rows")
