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
    ``outlier``: ``stress`` plus what trained BERT checkpoints show and random init does
    not: five hidden channels that carry 30x the magnitude of the rest in the residual
    stream (embedding columns, the rows of both residual-producing projections and their
    biases), LayerNorm gammas with outliers (4x on two of those channels, 0.25x on the other
    three, 0.05 on a few ordinary channels) and LayerNorm betas with a common offset, so
    that the residual stream has a non-zero mean.  As in trained models, the consumers
    have learned to discount the massive channels (their Q/K/FFN-up weight columns are
    30x smaller), which keeps the attention logits in a sane range: the model stays well
    conditioned (HF's own fp16 path reproduces its fp32 output to cosine > 0.9999), so the
    0.9995 bar is meaningful.  This is the case the folded LayerNorm (raw bf16 rows + fp32
    sum / sum-of-squares) has to survive.
    """
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}

    def normal(*size, std=0.02):
        return torch.empty(*size, dtype=torch.float32).normal_(0.0, std, generator=g)

    def uniform(n, lo, hi):
        return torch.empty(n, dtype=torch.float32).uniform_(lo, hi, generator=g)

    if style not in ("hf_init", "stress", "outlier"):
        raise ValueError(f"unknown weight style {style!r}")
    outlier = style == "outlier"
    stress = style == "stress" or outlier
    h, f = shape.hidden, shape.inter
    hot = [c % h for c in (7, 101, 250, 333, 380)]      # outlier channels
    cold = [c % h for c in (19, 64, 200, 301)]          # channels with a tiny gamma
    w["embeddings.word_embeddings.weight"] = normal(shape.vocab, h)
    if outlier:
        w["embeddings.word_embeddings.weight"][:, hot] *= 30.0
    w["embeddings.word_embeddings.weight"][0].zero_()  # padding_idx=0
    w["embeddings.position_embeddings.weight"] = normal(shape.max_pos, h)
    w["embeddings.token_type_embeddings.weight"] = normal(shape.type_vocab, h)

    def ln(prefix):
        w[prefix + ".weight"] = uniform(h, 0.5, 1.5) if stress else torch.ones(h)
        w[prefix + ".bias"] = uniform(h, -0.3, 0.3) if stress else torch.zeros(h)
        if outlier:
            w[prefix + ".weight"][hot[:2]] *= 4.0
            w[prefix + ".weight"][hot[2:]] *= 0.25
            w[prefix + ".weight"][cold] = 0.05
            w[prefix + ".bias"] += 0.4

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
            if outlier and lin in ("attention.output.dense", "output.dense"):
                w[p + lin + ".weight"][hot] *= 30.0
                w[p + lin + ".bias"][hot] += torch.tensor([2.0, -2.0, 1.5, -1.0, 2.5])
            if outlier and lin in ("attention.self.query", "attention.self.key", "intermediate.dense"):
                w[p + lin + ".weight"][:, hot] *= 1.0 / 30.0
        ln(p + "attention.output.LayerNorm")
        ln(p + "output.LayerNorm")
    if outlier:
        # the pooled embedding must not be carried by the massive channels alone (a cosine dominated by five
        # coordinates would hide errors in the other 379): the last LayerNorm tames them
        last = f"encoder.layer.{shape.layers - 1}.output.LayerNorm"
        w[last + ".weight"][hot] = 0.02
        w[last + ".bias"] = (w[last + ".bias"] - 0.4) * 0.1   # ... nor by a constant offset common to every input
    return w
