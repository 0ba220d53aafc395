"""Seeded synthetic BERT weights (HF state-dict naming) for the benchmarks and the parity tests.

Neutral input data: bench.py's CUDA arm, its CPU baseline arm and the oracle all draw the SAME
tensors from here, so that no measured path has to import anything under ``oracle/``
(torch.manual_seed-style N(0, 0.02) weights, zero biases, LayerNorm gamma=1 beta=0).
"""

from __future__ import annotations

from typing import Dict

import torch


class BgeSmallShape:
    """bge-small-en (BAAI/bge-small-en v1): the only architecture the reference loads (embeddings.py:28-32)."""
    vocab = 30522
    hidden = 384
    layers = 12
    heads = 12
    inter = 1536
    max_pos = 512
    type_vocab = 2
    ln_eps = 1e-12


def synth_weights(
    seed: int = 0, shape=BgeSmallShape, style: str = "hf_init"
) -> Dict[str, torch.Tensor]:
    """Seeded weights in HF naming.

    ``hf_init``: N(0, 0.02) matrices/embeddings, zero biases, LN gamma=1 beta=0
    (HF ``_init_weights``).  ``stress``: same plus non-zero biases, random LN
    affine and 6x larger Q/K projections so that attention is peaked and every
    bias/affine code path changes the result -- a harder numerical target.
    """
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}

    def normal(*size, std=0.02):
        return torch.empty(*size, dtype=torch.float32).normal_(0.0, std, generator=g)

    def uniform(n, lo, hi):
        return torch.empty(n, dtype=torch.float32).uniform_(lo, hi, generator=g)

    stress = style == "stress"
    h, f = shape.hidden, shape.inter
    w["embeddings.word_embeddings.weight"] = normal(shape.vocab, h)
    w["embeddings.word_embeddings.weight"][0].zero_()  # padding_idx=0
    w["embeddings.position_embeddings.weight"] = normal(shape.max_pos, h)
    w["embeddings.token_type_embeddings.weight"] = normal(shape.type_vocab, h)

    def ln(prefix):
        w[prefix + ".weight"] = uniform(h, 0.5, 1.5) if stress else torch.ones(h)
        w[prefix + ".bias"] = uniform(h, -0.3, 0.3) if stress else torch.zeros(h)

    ln("embeddings.LayerNorm")
    for i in range(shape.layers):
        p = f"encoder.layer.{i}."
        for lin, (o, k) in (
            ("attention.self.query", (h, h)),
            ("attention.self.key", (h, h)),
            ("attention.self.value", (h, h)),
            ("attention.output.dense", (h, h)),
            ("intermediate.dense", (f, h)),
            ("output.dense", (h, f)),
        ):
            std = 0.02
            if stress and lin in ("attention.self.query", "attention.self.key"):
                std = 0.12
            w[p + lin + ".weight"] = normal(o, k, std=std)
            w[p + lin + ".bias"] = uniform(o, -0.1, 0.1) if stress else torch.zeros(o)
        ln(p + "attention.output.LayerNorm")
        ln(p + "output.LayerNorm")
    return w
