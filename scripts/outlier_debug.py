"""Per-layer error of the CUDA encoder against the fp32 oracle for one weight style (default: outlier): where the
largest absolute errors sit (channel, reference value), per-token cosine.  python scripts/outlier_debug.py [style]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ai-dial-rag_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

from dial_rag_b200.embeddings.encoder import B200Encoder
from oracle import encoder as oenc
from tests.synth import synth_token_batch

style = sys.argv[1] if len(sys.argv) > 1 else "outlier"
seed = {"hf_init": 0, "stress": 7, "outlier": 11}[style]
w = oenc.synth_weights(seed=seed, style=style)
enc = B200Encoder(w, device=0, max_tokens=16384)
ids, cu = synth_token_batch(seed=31, n_seq=5, seq_len=200, ragged=True, min_len=5)
lists = oenc.packed_to_lists(ids, cu)
for layer in ((0, 1, 2, 6, 12) if len(sys.argv) <= 2 else (1, 2, 4, 6, 8, 10, 11, 12)):
    got = enc.debug_hidden(ids, cu, layer)
    worst = []
    for i, t in enumerate(lists[:2] if len(sys.argv) <= 2 else lists):
        tt = torch.tensor([t])
        ref = oenc.bert_hidden(w, tt, torch.ones_like(tt), oenc.BertShape(layers=layer))[0].numpy()
        g = got[cu[i]:cu[i + 1]]
        err = np.abs(g - ref)
        cos = (g * ref).sum(1) / (np.linalg.norm(g, axis=1) * np.linalg.norm(ref, axis=1))
        tok, ch = np.unravel_index(err.argmax(), err.shape)
        rel = err / (0.15 + 0.02 * np.abs(ref))
        rt, rc = np.unravel_index(rel.argmax(), rel.shape)
        if len(sys.argv) > 2:
            worst.append(f"seq {i} (len {len(t)}): min cos {cos.min():.6f} at token {int(cos.argmin())}, |ref| there {np.linalg.norm(ref[cos.argmin()]):.1f}, 2nd {np.sort(cos)[1]:.6f}")
            continue
        worst.append(f"seq {i}: min cos {cos.min():.6f}, max abs err {err.max():.3f} at ch {ch} (ref {ref[tok, ch]:.2f}), "
                     f"worst vs (0.15+2%) bound x{rel.max():.2f} at ch {rc} (ref {ref[rt, rc]:.3f}, got {g[rt, rc]:.3f}); "
                     f"err by channel top5 {np.argsort(-err.max(0))[:5].tolist()} {np.sort(err.max(0))[::-1][:5].round(3).tolist()}")
    print(f"layer {layer}: " + " | ".join(worst))
enc.close()
