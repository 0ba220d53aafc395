"""Phase timeline of the tcgen05 attention kernel (variant 7 = variant 3 + clock64 stamps of CTA 0).
Prints, per softmax warp, the median cycles of each phase of a key block and, per issuer, of its stages.
    python scripts/attn_trace.py [--seqs 1024 --len 256]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ai-dial-rag_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

from dial_rag_b200 import _native

ap = argparse.ArgumentParser()
ap.add_argument("--seqs", type=int, default=1024)
ap.add_argument("--len", type=int, default=256)
a = ap.parse_args()
lib = _native.load()
heads, hd = 12, 32
T = a.seqs * a.len
cu = torch.arange(0, T + 1, a.len, dtype=torch.int32, device="cuda")
g = torch.Generator(device="cuda").manual_seed(5)
qkv = (torch.randn(T, 3 * heads * hd, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
ctx = torch.empty((T, heads * hd), device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
words = lib.drag_debug_attention_trace_words()
trace = torch.zeros(words, dtype=torch.int64, device="cuda")
_native.check(lib.drag_debug_set_attention_trace(0, trace.data_ptr()))
for _ in range(2):
    trace.zero_()
    _native.check(lib.drag_debug_attention(0, 7, qkv.data_ptr(), ctx.data_ptr(), cu.data_ptr(), a.seqs, T, a.len, heads, st))
    torch.cuda.synchronize()
_native.check(lib.drag_debug_set_attention_trace(0, 0))
t = trace.cpu().numpy()
NB = 96
soft = t[: 8 * NB * 8].reshape(8, NB, 8)
iss = t[8 * NB * 8:].reshape(2, NB, 4)
names = ["wait s_full (job start only)", "ld S (4 x tcgen05.ld)", "row max", "exponentials", "wait S(next) / PV(prev)", "flush + rescale + st P", "fence + arrive", "-> next block top"]
print(f"{a.seqs} x {a.len}: cycles per phase of a key block (median over blocks 8..{NB - 8}), CTA 0")
for w in range(8):
    s = soft[w]
    ok = (s[:, 0] > 0) & (s[:, 7] > 0)
    idx = np.flatnonzero(ok)
    idx = idx[(idx >= 8) & (idx < NB - 8)]
    if len(idx) < 4:
        continue
    d = np.diff(s[idx], axis=1)
    nxt = s[idx[1:], 0] - s[idx[:-1], 7]
    total = s[idx[1:], 0] - s[idx[:-1], 0]
    print(f"warp {w} (group {w >> 2}): " + ", ".join(f"{n} {int(np.median(d[:, i]))}" for i, n in enumerate(names[:7])) +
          f", {names[7]} {int(np.median(nxt))} | block period {int(np.median(total))}")
for gidx in range(2):
    s = iss[gidx]
    idx = np.flatnonzero((s[:, 0] > 0) & (s[:, 2] > 0))
    idx = idx[(idx >= 8) & (idx < NB - 8)]
    if len(idx) < 4:
        continue
    pv_issue = np.median(s[idx, 1] - s[idx, 0])
    qk_issue = np.median(s[idx, 3] - s[idx, 2])
    period = np.median(np.diff(s[idx, 0]))
    print(f"issuer {gidx}: P V issue {int(pv_issue)} cycles, S issue {int(qk_issue)}, p_full -> next p_full {int(period)}; "
          f"after P V issued: s_free wait ends +{int(np.median(s[idx[1:], 2] - s[idx[:-1], 1]))}")
# phase offset between the groups: start of the exponentials of warp 4 (group 1) relative to warp 0 (group 0), same block ordinal
idx = np.arange(8, NB - 8)
off = soft[4][idx, 3] - soft[0][idx, 3]
print("group 1 starts its exponentials", int(np.median(off)), "cycles after group 0 (median; min", int(off.min()), "max", int(off.max()), ")")
# cross-role latencies for group 0, warp 0: p_full arrival -> P V issued; s_free arrival -> S issued
w0 = soft[0]
for gidx in range(1):
    s = iss[gidx]
    idx = np.arange(8, NB - 8)
    print("group 0: softmax arrive(p_full) -> issuer woke", int(np.median(s[idx, 0] - w0[idx, 7])),
          "; S loaded (s_free) -> issuer woke", int(np.median(s[idx + 1, 2] - w0[idx, 2])) if True else "")
