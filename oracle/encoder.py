"""CPU ORACLE (test infrastructure, not product code) -- bge-small-en encoder.

fp32 torch restatement of the embedding half of the hot path.  The arithmetic
of this half lives in third-party packages that are NOT under /root/reference:

  * langchain-community 0.3.20  ``HuggingFaceBgeEmbeddings``      (poetry.lock:2128)
  * sentence-transformers 3.3.1 ``SentenceTransformer.encode``    (poetry.lock:5943)
  * transformers 4.51.3         ``BertModel``                     (poetry.lock:6427)

so this file restates their published algorithm and is anchored on the
reference's call sites: aidial_rag/embeddings/embeddings.py:38-66 (model
construction, ``normalize_embeddings=True``), :79-96 (embed_documents /
embed_query) and :102-108 (batches of 128).

Parity status: PINNED against the ``transformers.BertModel`` that is installed
in the authoring container (5.5.0; same graph as 4.51.3 for BERT) on seeded
weights -- ``tests/golden/encoder_*.npz`` hold its outputs, produced by
``oracle/make_golden_encoder.py``; ``tests/test_oracle_encoder.py`` re-checks
the restatement against them and, when ``transformers`` is importable, against
a live ``BertModel``.  The reference's own tests hold NO numeric embedding
vectors, and the real bge-small-en weights/vocab are not available
offline, so real-weight parity is unpinned; the contract is BASELINE.md's
cosine >= 0.9995 against this fp32 oracle on identical token ids.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Sequence

import numpy as np
import torch

QUERY_INSTRUCTION = "Represent this question for searching relevant passages: "


@dataclass(frozen=True)
class BertShape:
    """bge-small-en (BAAI/bge-small-en v1 == epam/bge-small-en) config.json."""

    vocab: int = 30522
    hidden: int = 384
    layers: int = 12
    heads: int = 12
    inter: int = 1536
    max_pos: int = 512
    type_vocab: int = 2
    ln_eps: float = 1e-12

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


BGE_SMALL = BertShape()


def weight_names(shape: BertShape = BGE_SMALL) -> List[str]:
    """HF ``BertModel`` state-dict keys (no pooler: ST loads add_pooling_layer... the
    pooler output is never used by the CLS-pooling SentenceTransformer stack)."""
    names = [
        "embeddings.word_embeddings.weight",
        "embeddings.position_embeddings.weight",
        "embeddings.token_type_embeddings.weight",
        "embeddings.LayerNorm.weight",
        "embeddings.LayerNorm.bias",
    ]
    for i in range(shape.layers):
        p = f"encoder.layer.{i}."
        for lin in (
            "attention.self.query",
            "attention.self.key",
            "attention.self.value",
            "attention.output.dense",
            "intermediate.dense",
            "output.dense",
        ):
            names += [p + lin + ".weight", p + lin + ".bias"]
        names += [
            p + "attention.output.LayerNorm.weight",
            p + "attention.output.LayerNorm.bias",
            p + "output.LayerNorm.weight",
            p + "output.LayerNorm.bias",
        ]
    return names


def synth_weights(
    seed: int = 0, shape: BertShape = BGE_SMALL, style: str = "hf_init"
) -> Dict[str, torch.Tensor]:
    """Seeded weights in HF naming: the shared generator in ``synth_weights.py`` at the
    repo root (the benchmark's CUDA arm draws the same tensors from there without importing the oracle)."""
    import synth_weights as _sw

    return _sw.synth_weights(seed=seed, shape=shape, style=style)


def _layer_norm(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def _gelu_erf(x: torch.Tensor) -> torch.Tensor:
    # hidden_act="gelu" in bge-small-en == exact erf form
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


@torch.no_grad()
def bert_hidden(
    w: Dict[str, torch.Tensor],
    input_ids: torch.Tensor,
    attention_mask: torch.Tensor,
    shape: BertShape = BGE_SMALL,
) -> torch.Tensor:
    """``BertModel(input_ids, attention_mask).last_hidden_state`` in fp32.

    input_ids / attention_mask: int64 [B, S] (mask 1 = real token).  Follows
    BertEmbeddings (word + absolute position 0..S-1 + token type 0, LayerNorm),
    BertSelfAttention (softmax(QK^T/sqrt(d) + (1-mask)*min) V), BertSelfOutput,
    BertIntermediate (erf GELU), BertOutput.  Dropout is identity in eval.
    """
    B, S = input_ids.shape
    H, dh = shape.heads, shape.head_dim
    pos = torch.arange(S)
    x = (
        w["embeddings.word_embeddings.weight"][input_ids]
        + w["embeddings.position_embeddings.weight"][pos][None]
        + w["embeddings.token_type_embeddings.weight"][0][None, None]
    )
    x = _layer_norm(x, w["embeddings.LayerNorm.weight"], w["embeddings.LayerNorm.bias"], shape.ln_eps)
    key_bias = (1.0 - attention_mask.to(torch.float32))[:, None, None, :] * torch.finfo(torch.float32).min
    for i in range(shape.layers):
        p = f"encoder.layer.{i}."

        def lin(name, t):
            return t @ w[p + name + ".weight"].T + w[p + name + ".bias"]

        def heads(t):
            return t.view(B, S, H, dh).permute(0, 2, 1, 3)

        q, k, v = (heads(lin("attention.self." + n, x)) for n in ("query", "key", "value"))
        scores = q @ k.transpose(-1, -2) / math.sqrt(dh) + key_bias
        ctx = torch.softmax(scores, dim=-1) @ v
        ctx = ctx.permute(0, 2, 1, 3).reshape(B, S, H * dh)
        x = _layer_norm(
            lin("attention.output.dense", ctx) + x,
            w[p + "attention.output.LayerNorm.weight"],
            w[p + "attention.output.LayerNorm.bias"],
            shape.ln_eps,
        )
        inter = _gelu_erf(lin("intermediate.dense", x))
        x = _layer_norm(
            lin("output.dense", inter) + x,
            w[p + "output.LayerNorm.weight"],
            w[p + "output.LayerNorm.bias"],
            shape.ln_eps,
        )
    return x


@torch.no_grad()
def pool_and_normalize(hidden: torch.Tensor) -> torch.Tensor:
    """ST ``Pooling(cls)`` -> ``Normalize()`` module -> ``normalize_embeddings=True``.

    The bge-small-en SentenceTransformer stack L2-normalises twice (modules.json
    has a Normalize module and embeddings.py:60-62 asks encode() to normalise
    again); both use ``F.normalize(p=2, dim=1)`` (eps 1e-12).
    """
    cls = hidden[:, 0, :]
    once = torch.nn.functional.normalize(cls, p=2, dim=1)
    return torch.nn.functional.normalize(once, p=2, dim=1)


@torch.no_grad()
def encode_token_lists(
    w: Dict[str, torch.Tensor],
    token_lists: Sequence[Sequence[int]],
    shape: BertShape = BGE_SMALL,
    minibatch: int = 32,
    sort_key: Sequence[int] | None = None,
) -> np.ndarray:
    """``SentenceTransformer.encode`` on already-tokenised inputs -> f32 [n, hidden].

    Restates the ST 3.3.1 batching: order by length descending (ST sorts by
    *text* length; pass ``sort_key`` to reproduce that, default token count),
    minibatches of 32, pad with id 0 to the longest of the minibatch, truncate
    to ``max_pos`` keeping [CLS] first and [SEP] last, restore input order.
    """
    n = len(token_lists)
    out = np.zeros((n, shape.hidden), dtype=np.float32)
    toks = []
    for t in token_lists:
        t = list(t)
        if len(t) > shape.max_pos:
            t = t[: shape.max_pos - 1] + [t[-1]]
        toks.append(t)
    key = sort_key if sort_key is not None else [len(t) for t in toks]
    order = np.argsort([-int(k) for k in key], kind="stable")
    for s in range(0, n, minibatch):
        idx = order[s : s + minibatch]
        longest = max(len(toks[i]) for i in idx)
        ids = torch.zeros((len(idx), longest), dtype=torch.int64)
        mask = torch.zeros((len(idx), longest), dtype=torch.int64)
        for r, i in enumerate(idx):
            ids[r, : len(toks[i])] = torch.tensor(toks[i], dtype=torch.int64)
            mask[r, : len(toks[i])] = 1
        emb = pool_and_normalize(bert_hidden(w, ids, mask, shape))
        out[idx] = emb.numpy()
    return out


def packed_to_lists(ids: np.ndarray, cu_seqlens: np.ndarray) -> List[List[int]]:
    return [ids[cu_seqlens[i] : cu_seqlens[i + 1]].tolist() for i in range(len(cu_seqlens) - 1)]


def prepare_document_text(text: str) -> str:
    """HuggingFaceBgeEmbeddings.embed_documents: newlines -> spaces, no instruction."""
    return text.replace("\n", " ")


def prepare_query_text(text: str) -> str:
    """HuggingFaceBgeEmbeddings.embed_query: instruction prefix + text, newlines -> spaces."""
    return QUERY_INSTRUCTION + text.replace("\n", " ")


def hf_bert_model(w: Dict[str, torch.Tensor], shape: BertShape = BGE_SMALL):
    """A ``transformers.BertModel`` carrying exactly ``w`` (cross-check + CPU baseline)."""
    from transformers import BertConfig, BertModel

    cfg = BertConfig(
        vocab_size=shape.vocab,
        hidden_size=shape.hidden,
        num_hidden_layers=shape.layers,
        num_attention_heads=shape.heads,
        intermediate_size=shape.inter,
        max_position_embeddings=shape.max_pos,
        type_vocab_size=shape.type_vocab,
        layer_norm_eps=shape.ln_eps,
        hidden_act="gelu",
    )
    model = BertModel(cfg, add_pooling_layer=False)
    missing, unexpected = model.load_state_dict(w, strict=False)
    bad = [m for m in missing if "position_ids" not in m and "token_type_ids" not in m]
    if bad or unexpected:
        raise RuntimeError(f"state dict mismatch: missing={bad} unexpected={unexpected}")
    return model.eval()
