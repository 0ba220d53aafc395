"""Host logic of ``build_embeddings`` (aidial_rag/embeddings/embeddings.py:102-108 + batched.py:35-53) with a fake
encoder: ordered results over several batches, tokenise-ahead on the bounded indexing CPU pool, and a failing batch
that neither leaves a pending task behind nor an un-retrieved exception."""

import asyncio
import gc
import io
import threading
import warnings

import numpy as np
import pytest

from dial_rag_b200.embeddings import embeddings as emb


class FakeImpl:
    def __init__(self, fail_on_batch=None):
        self.fail_on_batch = fail_on_batch
        self.tokenized, self.embedded = [], []
        self.token_threads, self.embed_threads = set(), set()
        self.n_threads_seen = set()

    def tokenize_documents(self, texts, n_threads=0):
        self.token_threads.add(threading.current_thread().name)
        self.n_threads_seen.add(n_threads)
        self.tokenized.append(list(texts))
        ids = np.array([int(t) for t in texts], dtype=np.int32)
        return ids, np.arange(len(texts) + 1, dtype=np.int32)

    def embed_packed_numpy(self, ids, cu):
        self.embed_threads.add(threading.current_thread().name)
        if self.fail_on_batch is not None and len(self.embedded) == self.fail_on_batch:
            raise RuntimeError("device lost")
        self.embedded.append(ids.copy())
        return np.repeat(ids[:, None].astype(np.float32), 4, axis=1)


@pytest.fixture
def small_batches(monkeypatch):
    monkeypatch.setattr(emb, "EMBEDDINGS_BATCH_SIZE", 3)
    yield
    emb.configure(None)


def test_ordered_results_one_batch_at_a_time_and_progress_lines(small_batches):
    fake = FakeImpl()
    emb.configure(fake)
    stage = io.StringIO()
    rows = list(asyncio.run(emb.build_embeddings((str(i) for i in range(10)), stage)))
    assert [int(r[0]) for r in rows] == list(range(10)) and rows[0].dtype == np.float32
    assert [len(b) for b in fake.tokenized] == [3, 3, 3, 1]
    assert fake.n_threads_seen == {emb.TOKENIZER_THREADS} and 1 <= emb.TOKENIZER_THREADS <= 8
    assert all(n.startswith("indexing_cpu") for n in fake.token_threads), fake.token_threads
    assert all(n.startswith("indexing_embeddings") for n in fake.embed_threads), fake.embed_threads
    assert "4/4" in stage.getvalue()        # tqdm keep-alive lines, one bar over the batches


def test_empty_input(small_batches):
    emb.configure(FakeImpl())
    assert list(asyncio.run(emb.build_embeddings(iter(()), io.StringIO()))) == []


def test_failing_batch_propagates_and_leaves_nothing_pending(small_batches):
    fake = FakeImpl(fail_on_batch=1)
    emb.configure(fake)
    leftovers = []

    async def run():
        try:
            await emb.build_embeddings((str(i) for i in range(12)), io.StringIO())
        finally:
            leftovers.extend(t for t in asyncio.all_tasks() if t is not asyncio.current_task() and not t.done())

    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        with pytest.raises(RuntimeError, match="device lost"):
            asyncio.run(run())
        gc.collect()
    assert not leftovers
    assert not [w for w in caught if "never retrieved" in str(w.message) or "was never awaited" in str(w.message)]
    assert len(fake.embedded) == 1 and len(fake.tokenized) <= 3   # at most one batch was tokenised ahead of the failure
