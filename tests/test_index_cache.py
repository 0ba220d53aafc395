"""Host logic of the cross-request resident-index cache: keyed by VALUE (the build id this package
stores in the first item's persisted ``id`` field, else a content digest), so records deserialised afresh from the
same stored bytes -- what the reference's storage hands out on every request (index_storage.py:56-66, :136) --
hit; LRU by bytes and entries; no weak references (nothing can re-enter the lock)."""

import gzip
import pickle
import threading

import numpy as np

from dial_rag_b200.records import ItemEmbeddings, MultiEmbeddings
from dial_rag_b200.retrievers.embeddings_index import (
    UID_PREFIX,
    ResidentIndexCache,
    pack_embedding_matrix,
    pack_simple_embeddings,
    source_key,
)


def _rows(seed, n, dim=8):
    return np.random.default_rng(seed).standard_normal((n, dim)).astype(np.float32)


def _docs():
    return pack_simple_embeddings(_rows(1, 2)), pack_simple_embeddings(_rows(2, 1)), pack_simple_embeddings(_rows(3, 3))


def _roundtrip(multi):
    """The reference's persisted form: DocumentRecord.to_bytes(protocol='pickle', compress='gzip') (index_storage.py:44)."""
    return pickle.loads(gzip.decompress(gzip.compress(pickle.dumps(multi))))


def test_hit_needs_the_same_documents_in_the_same_order():
    a, b, c = _docs()
    cache = ResidentIndexCache(max_entries=4, max_bytes=1 << 20)
    assert cache.get([a, b], ("f32", None)) is None
    assert cache.put([a, b], ("f32", None), "AB", 100)
    assert cache.get([a, b], ("f32", None)) == "AB"
    assert cache.get([b, a], ("f32", None)) is None          # document order defines row ids
    assert cache.get([a, b], ("bf16", None)) is None         # storage / device are part of the key
    assert cache.get([a, b, c], ("f32", None)) is None
    assert cache.hits == 1 and cache.misses == 4


def test_deserialised_copies_hit():
    """fresh objects per request (from_bytes on every load) must reuse the resident matrix."""
    a, b, _ = _docs()
    cache = ResidentIndexCache()
    cache.put([a, b], ("f32", 0), "AB", 1)
    a2, b2 = _roundtrip(a), _roundtrip(b)
    assert a2 is not a and a2[0] is not a[0]
    assert a2[0].id == a[0].id and a2[0].id.startswith(UID_PREFIX)
    assert cache.get([a2, b2], ("f32", 0)) == "AB"
    assert cache.get([_roundtrip(a2), _roundtrip(b2)], ("f32", 0)) == "AB"


def test_indexes_without_build_id_are_keyed_by_content():
    """Records written by the stock reference carry no id: the digest of their rows is the key."""
    rows = _rows(5, 4)
    plain = MultiEmbeddings([ItemEmbeddings(embeddings=rows[i:i + 1].copy()) for i in range(4)])
    assert plain[0].id is None and source_key(plain)[0] == "content"
    cache = ResidentIndexCache()
    cache.put([plain], (), "P", 1)
    assert cache.get([_roundtrip(plain)], ()) == "P"
    other = MultiEmbeddings([ItemEmbeddings(embeddings=rows[i:i + 1] + (i == 2)) for i in range(4)])
    assert cache.get([other], ()) is None                     # one changed row in the middle: different digest


def test_same_rows_built_twice_do_not_alias_by_id_but_changed_ends_miss():
    rows = _rows(6, 5)
    a, b = pack_simple_embeddings(rows), pack_simple_embeddings(rows)
    assert a[0].id != b[0].id                                 # every build gets its own id
    cache = ResidentIndexCache()
    cache.put([a], (), "A", 1)
    assert cache.get([b], ()) is None
    grown = MultiEmbeddings(list(a) + [ItemEmbeddings(embeddings=rows[:1])])
    assert cache.get([grown], ()) is None                     # a re-indexed document grew: the count is in the key
    a[-1].embeddings = a[-1].embeddings + 1.0
    assert cache.get([a], ()) is None                         # ... or its boundary rows changed


def test_lru_eviction_by_bytes_and_entries():
    a, b, c = _docs()
    cache = ResidentIndexCache(max_entries=2, max_bytes=100)
    cache.put([a], (), "A", 60)
    cache.put([b], (), "B", 30)
    assert cache.get([a], ()) == "A"                          # A is now the most recent
    cache.put([c], (), "C", 30)                               # 120 bytes: the least recent (B) goes
    assert cache.get([b], ()) is None and cache.get([a], ()) == "A" and cache.get([c], ()) == "C"
    big = pack_simple_embeddings(_rows(9, 1))
    cache.put([big], (), "BIG", 1000)                         # a single oversized entry is still kept
    assert cache.get([big], ()) == "BIG" and len(cache) == 1


def test_entry_size_is_asked_at_eviction_time():
    """An entry whose footprint grows after insertion (the bf16 scoring copy of the batched path) counts in full."""
    a, b, _ = _docs()
    size = {"a": 10}
    cache = ResidentIndexCache(max_entries=8, max_bytes=100)
    cache.put([a], (), "A", lambda: size["a"])
    size["a"] = 95
    assert cache.nbytes() == 95
    cache.put([b], (), "B", 10)                               # 105 bytes now: A (least recent) goes
    assert cache.get([a], ()) is None and cache.get([b], ()) == "B"


def test_concurrent_get_put_do_not_deadlock():
    docs = [pack_simple_embeddings(_rows(100 + i, 3)) for i in range(8)]
    cache = ResidentIndexCache(max_entries=4, max_bytes=1 << 20)

    def worker(k):
        for i in range(200):
            d = docs[(k + i) % len(docs)]
            if cache.get([d], ()) is None:
                cache.put([d], (), k, 10)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=30)
    assert not any(t.is_alive() for t in threads) and len(cache) <= 4


def test_contiguous_matrix_pack_is_zero_copy_and_flattens_in_o1():
    from dial_rag_b200.retrievers.embeddings_index import create_index_by_chunk

    m = _rows(7, 1000, 16)
    multi = pack_embedding_matrix(m)
    assert len(multi) == 1000 and multi[3].embeddings.shape == (1, 16)
    assert np.shares_memory(np.asarray(multi[3].embeddings), m)
    flat = create_index_by_chunk(multi)
    assert flat.embeddings is m or np.shares_memory(flat.embeddings, m)
    assert np.array_equal(flat.chunk_ids, np.arange(1000)) and flat.chunk_ids.dtype == np.int64
    # after a round trip the shortcut is gone, the result is not
    again = create_index_by_chunk(_roundtrip(multi))
    assert np.array_equal(again.embeddings, m) and np.array_equal(again.chunk_ids, flat.chunk_ids)
    # a container that no longer is n ordered views falls back to the general path
    multi[500] = ItemEmbeddings(embeddings=np.zeros((2, 16), dtype=np.float32))
    slow = create_index_by_chunk(multi)
    assert len(slow.embeddings) == 1001 and slow.chunk_ids[500] == 500 and slow.chunk_ids[501] == 500
