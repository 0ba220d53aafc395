"""The real-checkpoint constructor path (aidial_rag/embeddings/embeddings.py:28-36, :52-66):
``BGE_EMBEDDINGS_MODEL_PATH`` -> ``bge_embedding_impl()`` -> ``B200BgeEmbeddings.from_model_dir`` ->
``load_model_dir`` (model.safetensors, with or without the ``bert.`` prefix) + ``WordPieceTokenizer.from_model_dir``
(tokenizer.json / vocab.txt).  The real bge-small-en files are not available offline, so the directory is written
here in the checkpoint's own format (HF state-dict names, safetensors, the stock uncased BERT tokenizer.json) with
the seeded weights of the parity tests.
"""

import importlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import encoder as oenc
from tests.text_fixture import CHUNKS, QUERIES, build_vocab_file


def write_checkpoint_dir(path: str, weights, prefix: str = "", with_tokenizer_json: bool = True, extra_pooler: bool = True) -> str:
    """An HF BERT checkpoint directory: config.json, model.safetensors, vocab.txt (+ tokenizer.json)."""
    from safetensors.numpy import save_file

    os.makedirs(path, exist_ok=True)
    tensors = {prefix + k: np.ascontiguousarray(v.numpy()) for k, v in weights.items()}
    # what real checkpoints carry besides the encoder: position ids and a pooler the reference never uses
    tensors[prefix + "embeddings.position_ids"] = np.arange(512, dtype=np.int64)[None, :]
    if extra_pooler:
        tensors[prefix + "pooler.dense.weight"] = np.zeros((384, 384), dtype=np.float32)
        tensors[prefix + "pooler.dense.bias"] = np.zeros(384, dtype=np.float32)
    save_file(tensors, os.path.join(path, "model.safetensors"))
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump({"architectures": ["BertModel"], "hidden_size": 384, "num_hidden_layers": 12, "num_attention_heads": 12,
                   "intermediate_size": 1536, "vocab_size": 30522, "max_position_embeddings": 512, "type_vocab_size": 2,
                   "layer_norm_eps": 1e-12, "hidden_act": "gelu", "model_type": "bert"}, f)
    vocab = build_vocab_file(path)
    if with_tokenizer_json:
        from tokenizers import BertWordPieceTokenizer

        BertWordPieceTokenizer(vocab, lowercase=True).save(os.path.join(path, "tokenizer.json"))
    return path


@pytest.fixture(scope="module")
def weights():
    return oenc.synth_weights(seed=7, style="stress")


@pytest.mark.parametrize("prefix", ["", "bert."])
def test_load_model_dir_reads_safetensors(tmp_path, weights, prefix):
    """CPU: every tensor drag_encoder_create needs comes back under its un-prefixed HF name, bit for bit."""
    from dial_rag_b200.embeddings.encoder import load_model_dir, weight_order

    d = write_checkpoint_dir(str(tmp_path / "ckpt"), weights, prefix=prefix)
    got = load_model_dir(d)
    for name in weight_order():
        assert name in got, name
        assert got[name].dtype == np.float32
        assert np.array_equal(got[name], weights[name].numpy()), name


def test_load_model_dir_reads_pytorch_bin(tmp_path, weights):
    from dial_rag_b200.embeddings.encoder import load_model_dir, weight_order

    d = str(tmp_path / "ckpt")
    os.makedirs(d)
    torch.save({"bert." + k: v.half() for k, v in weights.items()}, os.path.join(d, "pytorch_model.bin"))
    got = load_model_dir(d)
    name = weight_order()[7]
    assert got[name].dtype == np.float32 and np.array_equal(got[name], weights[name].half().float().numpy())


@pytest.mark.parametrize("with_json", [True, False], ids=["tokenizer.json", "vocab.txt"])
def test_tokenizer_from_model_dir_matches_reference_tokenizer(tmp_path, weights, with_json):
    """CPU: the tokenizer loaded from the directory gives the ids of HF's BertWordPieceTokenizer on the same vocab."""
    from tokenizers import BertWordPieceTokenizer

    from dial_rag_b200.embeddings.tokenizer import WordPieceTokenizer

    d = write_checkpoint_dir(str(tmp_path / "ckpt"), weights, with_tokenizer_json=with_json)
    tok = WordPieceTokenizer.from_model_dir(d)
    ref = BertWordPieceTokenizer(os.path.join(d, "vocab.txt"), lowercase=True)
    ref.enable_truncation(max_length=512)
    texts = CHUNKS[:5] + QUERIES + ["Über-naïve café\n\nnew line", "x" * 3000]
    assert tok.encode_batch(texts) == [e.ids for e in ref.encode_batch(texts)]
    ids, cu = tok.encode_packed(texts)
    assert ids.tolist() == [t for e in ref.encode_batch(texts) for t in e.ids]
    assert np.diff(cu).max() <= 512


def test_missing_directory_raises(monkeypatch, tmp_path):
    from dial_rag_b200.embeddings import embeddings as emb

    monkeypatch.setattr(emb, "BGE_EMBEDDINGS_MODEL_NAME_OR_PATH", str(tmp_path / "nope"))
    monkeypatch.setattr(emb, "BGE_EMBEDDINGS_DEVICE", "cuda")
    emb.configure(None)
    with pytest.raises((FileNotFoundError, RuntimeError)):
        emb.bge_embedding_impl()


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("prefix", ["", "bert."])
def test_env_path_to_embed_query_on_gpu(tmp_path, weights, prefix, monkeypatch):
    """GPU: BGE_EMBEDDINGS_MODEL_PATH=<dir> -> module import -> bge_embedding_impl() -> embed_query / embed_documents /
    aembed_query, against the fp32 oracle on the same text preparation and ids (cosine >= 0.9995)."""
    import asyncio

    d = write_checkpoint_dir(str(tmp_path / "ckpt"), weights, prefix=prefix)
    monkeypatch.setenv("BGE_EMBEDDINGS_MODEL_PATH", d)
    monkeypatch.setenv("BGE_EMBEDDINGS_DEVICE", "cuda")
    from dial_rag_b200.embeddings import embeddings as emb

    emb.configure(None)
    emb = importlib.reload(emb)   # the module reads the environment at import, like the reference (embeddings.py:28-36)
    try:
        assert emb.BGE_EMBEDDINGS_MODEL_NAME_OR_PATH == d
        impl = emb.bge_embedding_impl()
        assert impl is emb.bge_embedding_impl()   # created once (reference: @cache)
        tok = impl.tokenizer
        q = QUERIES[0]
        got_q = np.array(emb.bge_embedding.embed_query(q))
        assert got_q.shape == (384,) and got_q.dtype == np.float64
        want_q = oenc.encode_token_lists(weights, tok.encode_batch([oenc.prepare_query_text(q)]))[0]
        assert float(got_q @ want_q) >= 0.9995
        got_async = np.array(asyncio.run(emb.bge_embedding.aembed_query(q)))
        assert np.array_equal(got_async, got_q)
        docs = CHUNKS[:9]
        got_d = np.array(emb.bge_embedding.embed_documents(docs))
        want_d = oenc.encode_token_lists(weights, tok.encode_batch([oenc.prepare_document_text(t) for t in docs]))
        assert ((got_d * want_d).sum(1) >= 0.9995).all()
        rows = asyncio.run(emb.bge_embedding.aembed_documents_numpy(docs))
        assert len(rows) == 9 and rows[0].dtype == np.float32 and np.allclose(np.stack(rows), got_d, atol=1e-7)
    finally:
        if emb._impl is not None:
            emb._impl.client.close()
        emb.configure(None)
        monkeypatch.delenv("BGE_EMBEDDINGS_MODEL_PATH")
        importlib.reload(emb)
