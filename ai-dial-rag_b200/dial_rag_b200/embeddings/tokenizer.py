"""WordPiece tokenisation for the encoder (host side, HF ``tokenizers``).

The reference tokenises inside sentence-transformers with the checkpoint's
``BertTokenizerFast`` (uncased WordPiece, [CLS]=101 ... [SEP]=102, truncation at 512;
SURVEY 8a row a4).  The same Rust tokenizer is used here, loaded from the checkpoint
directory (``tokenizer.json`` or ``vocab.txt``); it stays on the host by design
 and feeds packed int32 ids to the CUDA encoder.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np


class WordPieceTokenizer:
    """The checkpoint's uncased WordPiece tokenizer.  ``encode_packed`` is the encoder's feed: ASCII texts
    go through the multi-threaded C++ fast path of the C-ABI library (``drag_wordpiece_encode``: the same
    clean-up / lower-casing / punctuation split / greedy WordPiece / [CLS]..[SEP] template), everything
    else (non-ASCII text, literal special tokens) through the reference Rust tokenizer, text by text."""

    def __init__(self, tokenizer, max_length: int = 512, vocab_path: Optional[str] = None, lowercase: bool = True):
        self._tok = tokenizer
        self.max_length = max_length
        self._tok.enable_truncation(max_length=max_length)
        self._tok.no_padding()
        self._native = None      # drag_wordpiece handle (created on first use: needs the built library)
        self._vocab_path = vocab_path
        self._lowercase = lowercase

    def _fast(self):
        if self._native is None and self._vocab_path is not None:
            from dial_rag_b200 import _native

            lib = _native.load()
            handle = C.c_void_p()
            _native.check(lib.drag_wordpiece_create(self._vocab_path.encode(), int(self._lowercase), C.byref(handle)))
            self._native = (lib, handle)
        return self._native

    def __del__(self):  # pragma: no cover
        try:
            if self._native is not None:
                self._native[0].drag_wordpiece_destroy(self._native[1])
        except Exception:  # noqa: BLE001
            pass

    @classmethod
    def from_model_dir(cls, path: str, max_length: int = 512) -> "WordPieceTokenizer":
        from tokenizers import BertWordPieceTokenizer, Tokenizer

        tj = os.path.join(path, "tokenizer.json")
        vocab = os.path.join(path, "vocab.txt")
        if os.path.exists(tj):
            # the fast path restates the stock uncased BERT pipeline: only used next to a plain vocab.txt
            return cls(Tokenizer.from_file(tj), max_length, vocab_path=vocab if os.path.exists(vocab) else None)
        if os.path.exists(vocab):
            return cls.from_vocab_file(vocab, max_length)
        raise FileNotFoundError(f"no tokenizer.json / vocab.txt under {path}")

    @classmethod
    def from_vocab_file(cls, vocab_path: str, max_length: int = 512, lowercase: bool = True) -> "WordPieceTokenizer":
        from tokenizers import BertWordPieceTokenizer

        return cls(BertWordPieceTokenizer(vocab_path, lowercase=lowercase), max_length, vocab_path=vocab_path, lowercase=lowercase)

    def encode_batch(self, texts: Sequence[str]) -> List[List[int]]:
        if not texts:
            return []
        return [e.ids for e in self._tok.encode_batch(list(texts))]

    def encode_packed(self, texts: Sequence[str], n_threads: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """``(ids int32[T], cu_seqlens int32[n+1])`` of the packed batch -- what the CUDA encoder consumes."""
        n = len(texts)
        fast = self._fast() if n else None
        if fast is None:
            lists = self.encode_batch(texts)
            lens = np.fromiter((len(t) for t in lists), dtype=np.int64, count=n)
            cu = np.zeros(n + 1, dtype=np.int32)
            np.cumsum(lens, out=cu[1:])
            ids = np.fromiter((t for ids_ in lists for t in ids_), dtype=np.int32, count=int(cu[-1]))
            return ids, cu
        lib, handle = fast
        from dial_rag_b200 import _native

        raw = [t.encode("utf-8") for t in texts]
        offsets = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.fromiter((len(b) for b in raw), dtype=np.int64, count=n), out=offsets[1:])
        blob = b"".join(raw)
        out_ids = np.empty((n, self.max_length), dtype=np.int32)
        out_len = np.empty(n, dtype=np.int32)
        _native.check(lib.drag_wordpiece_encode(handle, blob, offsets.ctypes.data, n, self.max_length, int(n_threads),
                                                out_ids.ctypes.data, out_len.ctypes.data))
        slow = np.flatnonzero(out_len < 0)
        if len(slow):
            for i, enc in zip(slow, self._tok.encode_batch([texts[j] for j in slow]), strict=True):
                out_len[i] = len(enc.ids)
                out_ids[i, : out_len[i]] = enc.ids
        cu = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(out_len, out=cu[1:])
        ids = out_ids[np.arange(self.max_length)[None, :] < out_len[:, None]]   # row-major: the packed order
        return np.ascontiguousarray(ids, dtype=np.int32), cu
