"""Generate tests/golden/encoder_*.npz with the installed ``transformers.BertModel``.

Run in the authoring container:   python oracle/make_golden_encoder.py

The embedding arithmetic of the reference is third-party (see oracle/encoder.py
header); the closest thing to "the reference itself" that can execute offline
is HF ``BertModel`` (the reference's backend="torch" path,
aidial_rag/embeddings/embeddings.py:43-48) in fp32 with seeded weights.  Two
fixtures are written:

  * encoder_tiny.npz   -- a 2-layer/64-d BERT whose *weights are stored in the
    file*, so the check does not depend on the torch RNG;
  * encoder_bge.npz    -- the full bge-small-en shape, seeded weights
    (``synth_weights(seed, style)``) identified by a sha256 of the tensors,
    ragged sequences up to 512 tokens; outputs of HF BertModel + CLS pooling +
    double L2 normalisation.
"""

from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import encoder as enc  # noqa: E402
from tests.synth import synth_token_batch  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def weights_digest(w) -> str:
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode())
        h.update(w[k].numpy().tobytes())
    return h.hexdigest()


@torch.no_grad()
def hf_embed(model, token_lists, pad_to=None):
    out = []
    for t in token_lists:
        ids = torch.tensor([t], dtype=torch.int64)
        mask = torch.ones_like(ids)
        if pad_to:
            ids = torch.nn.functional.pad(ids, (0, pad_to - len(t)))
            mask = torch.nn.functional.pad(mask, (0, pad_to - len(t)))
        hidden = model(input_ids=ids, attention_mask=mask).last_hidden_state
        out.append(enc.pool_and_normalize(hidden)[0].numpy())
    return np.stack(out)


def main() -> None:
    torch.set_num_threads(8)
    os.makedirs(GOLDEN, exist_ok=True)

    # ---- tiny model, weights stored
    tiny = enc.BertShape(vocab=120, hidden=64, layers=2, heads=4, inter=128, max_pos=64)
    w = enc.synth_weights(seed=3, shape=tiny, style="stress")
    model = enc.hf_bert_model(w, tiny)
    rng = np.random.Generator(np.random.PCG64(5))
    lens = [1, 2, 9, 33, 64]
    toks = [rng.integers(1, tiny.vocab, size=n).tolist() for n in lens]
    emb = hf_embed(model, toks)
    emb_padded = hf_embed(model, toks, pad_to=64)
    np.savez_compressed(
        os.path.join(GOLDEN, "encoder_tiny.npz"),
        embeddings=emb,
        embeddings_padded=emb_padded,
        lens=np.array(lens),
        tokens=np.concatenate([np.array(t) for t in toks]),
        **{"w:" + k: v.numpy() for k, v in w.items()},
    )

    # ---- full bge-small-en shape, seeded weights
    recs = {}
    for style, seed in (("hf_init", 0), ("stress", 7), ("outlier", 11)):
        w = enc.synth_weights(seed=seed, style=style)
        model = enc.hf_bert_model(w)
        ids, cu = synth_token_batch(seed=21, n_seq=10, seq_len=512, ragged=True, min_len=3)
        # force a few exact lengths: 512 (max), 256 (bench shape), 2 (CLS+SEP only)
        lens = np.diff(cu).tolist()
        toks = enc.packed_to_lists(ids, cu)
        toks[0] = (toks[0] * 200)[:511] + [102]
        toks[0][0] = 101
        toks[1] = (toks[1] * 200)[:255] + [102]
        toks[1][0] = 101
        toks[2] = [101, 102]
        emb = hf_embed(model, toks)
        recs[f"{style}:seed"] = np.array(seed)
        recs[f"{style}:digest"] = np.array(weights_digest(w))
        recs[f"{style}:embeddings"] = emb
        recs[f"{style}:lens"] = np.array([len(t) for t in toks])
        recs[f"{style}:tokens"] = np.concatenate([np.array(t) for t in toks])
        print(style, "lens", [len(t) for t in toks], "digest", recs[f"{style}:digest"])
    np.savez_compressed(os.path.join(GOLDEN, "encoder_bge.npz"), **recs)
    print("wrote", sorted(os.listdir(GOLDEN)))


if __name__ == "__main__":
    main()
