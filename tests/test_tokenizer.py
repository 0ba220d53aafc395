"""The C++ WordPiece fast path (drag_wordpiece_encode) against the reference tokenizer
(HF `tokenizers` BertWordPieceTokenizer == transformers BertTokenizerFast): identical ids
on every text, with non-ASCII texts and literal special tokens routed through the reference tokenizer."""

import random
import string

import numpy as np
import pytest

from dial_rag_b200.embeddings.tokenizer import WordPieceTokenizer
from tests.text_fixture import CHUNKS, QUERIES, build_vocab_file


@pytest.fixture(scope="module")
def tok(tmp_path_factory):
    return WordPieceTokenizer.from_vocab_file(build_vocab_file(str(tmp_path_factory.mktemp("vocab"))))


def _check(tok, texts, **kw):
    ids, cu = tok.encode_packed(texts, **kw)
    want = tok.encode_batch(texts)
    assert ids.dtype == np.int32 and cu.dtype == np.int32 and cu[0] == 0 and len(cu) == len(texts) + 1
    assert [int(x) for x in np.diff(cu)] == [len(w) for w in want]
    for i, w in enumerate(want):
        assert ids[cu[i]:cu[i + 1]].tolist() == w, repr(texts[i])[:80]


def test_corpus_and_queries(tok):
    _check(tok, CHUNKS + QUERIES + [c.replace("\n", " ") for c in CHUNKS])


def test_truncation_to_512_keeps_cls_and_sep(tok):
    long_texts = [(c + " ") * 12 for c in CHUNKS[:8]]          # > 512 tokens each
    ids, cu = tok.encode_packed(long_texts)
    assert set(np.diff(cu).tolist()) == {512}
    assert all(ids[cu[i]] == 101 and ids[cu[i + 1] - 1] == 102 for i in range(len(long_texts)))
    _check(tok, long_texts)


def test_random_ascii_including_control_characters_and_punctuation(tok):
    rng = random.Random(1234)
    alphabet = string.ascii_letters + string.digits + string.punctuation + "    \t\n\r\x00\x07\x1f\x7f"
    texts = ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, 400))) for _ in range(400)]
    _check(tok, texts)


def test_edge_cases(tok):
    _check(tok, ["", " ", "\n\t", "a", "A", "a" * 100, "a" * 101, "glacier" * 30, "don't stop-me,now!", "UPPER lower MiXeD",
                 "x" * 99 + "." + "y" * 120, "...", "word##piece", "## ##a"])


def test_non_ascii_and_literal_special_tokens_take_the_reference_path(tok):
    texts = ["café naïve résumé", "中文 and english", "emoji \U0001f600 inside", "has [SEP] inside", "[CLS]",
             "x [MASK] y [PAD] [UNK]", "plain ascii between", " nbsp and – dash “quotes”", "[sep] lower-case is not special"]
    _check(tok, texts)


@pytest.mark.parametrize("threads", [1, 2, 3, 0])
def test_thread_counts_give_the_same_packing(tok, threads):
    _check(tok, (CHUNKS + QUERIES) * 3, n_threads=threads)


def test_empty_batch(tok):
    ids, cu = tok.encode_packed([])
    assert len(ids) == 0 and cu.tolist() == [0]
