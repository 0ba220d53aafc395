"""Host logic of the cross-request resident-index cache: keyed by the identity of the
documents' persisted MultiEmbeddings objects, weakly referenced, LRU by bytes and entries."""

import gc

from dial_rag_b200.records import MultiEmbeddings
from dial_rag_b200.retrievers.embeddings_index import ResidentIndexCache


def _docs():
    return MultiEmbeddings([1, 2]), MultiEmbeddings([3]), MultiEmbeddings([4, 5, 6])


def test_hit_needs_the_same_objects_in_the_same_order():
    a, b, c = _docs()
    cache = ResidentIndexCache(max_entries=4, max_bytes=1 << 20)
    assert cache.get([a, b], ("f32", None)) is None
    assert cache.put([a, b], ("f32", None), "AB", 100)
    assert cache.get([a, b], ("f32", None)) == "AB"
    assert cache.get([b, a], ("f32", None)) is None          # document order defines row ids
    assert cache.get([a, b], ("bf16", None)) is None         # storage / device are part of the key
    assert cache.get([a, b, c], ("f32", None)) is None
    assert cache.hits == 1 and cache.misses == 4


def test_equal_but_distinct_objects_do_not_alias():
    a, _, _ = _docs()
    twin = MultiEmbeddings(list(a))
    cache = ResidentIndexCache()
    cache.put([a], (), "A", 1)
    assert cache.get([twin], ()) is None and cache.get([a], ()) == "A"


def test_mutated_source_invalidates_the_entry():
    a, _, _ = _docs()
    cache = ResidentIndexCache()
    cache.put([a], (), "A", 1)
    a.append(7)                                               # a re-indexed document grew
    assert cache.get([a], ()) is None and len(cache) == 0


def test_collected_source_drops_the_entry():
    a, b, _ = _docs()
    cache = ResidentIndexCache()
    cache.put([a, b], (), "AB", 1)
    del b
    gc.collect()
    assert len(cache) == 0                                    # the HBM the entry pinned is released


def test_lru_eviction_by_bytes_and_entries():
    a, b, c = _docs()
    cache = ResidentIndexCache(max_entries=2, max_bytes=100)
    cache.put([a], (), "A", 60)
    cache.put([b], (), "B", 30)
    assert cache.get([a], ()) == "A"                          # A is now the most recent
    cache.put([c], (), "C", 30)                               # 120 bytes: the least recent (B) goes
    assert cache.get([b], ()) is None and cache.get([a], ()) == "A" and cache.get([c], ()) == "C"
    big = MultiEmbeddings([0])
    cache.put([big], (), "BIG", 1000)                         # a single oversized entry is still kept
    assert cache.get([big], ()) == "BIG" and len(cache) == 1


def test_unweakrefable_sources_are_not_cached():
    cache = ResidentIndexCache()
    assert not cache.put([[1, 2]], (), "X", 1)                # plain lists cannot be weakly referenced
    assert cache.get([[1, 2]], ()) is None
