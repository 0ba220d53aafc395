"""WordPiece tokenisation for the encoder (host side, HF ``tokenizers``).

The reference tokenises inside sentence-transformers with the checkpoint's
``BertTokenizerFast`` (uncased WordPiece, [CLS]=101 ... [SEP]=102, truncation at 512;
SURVEY 8a row a4).  The same Rust tokenizer is used here, loaded from the checkpoint
directory (``tokenizer.json`` or ``vocab.txt``); it stays on the host by design
 and feeds packed int32 ids to the CUDA encoder.
"""

from __future__ import annotations

import os
from typing import List, Sequence


class WordPieceTokenizer:
    def __init__(self, tokenizer, max_length: int = 512):
        self._tok = tokenizer
        self.max_length = max_length
        self._tok.enable_truncation(max_length=max_length)
        self._tok.no_padding()

    @classmethod
    def from_model_dir(cls, path: str, max_length: int = 512) -> "WordPieceTokenizer":
        from tokenizers import BertWordPieceTokenizer, Tokenizer

        tj = os.path.join(path, "tokenizer.json")
        if os.path.exists(tj):
            return cls(Tokenizer.from_file(tj), max_length)
        vocab = os.path.join(path, "vocab.txt")
        if os.path.exists(vocab):
            return cls.from_vocab_file(vocab, max_length)
        raise FileNotFoundError(f"no tokenizer.json / vocab.txt under {path}")

    @classmethod
    def from_vocab_file(cls, vocab_path: str, max_length: int = 512, lowercase: bool = True) -> "WordPieceTokenizer":
        from tokenizers import BertWordPieceTokenizer

        return cls(BertWordPieceTokenizer(vocab_path, lowercase=lowercase), max_length)

    def encode_batch(self, texts: Sequence[str]) -> List[List[int]]:
        if not texts:
            return []
        return [e.ids for e in self._tok.encode_batch(list(texts))]
