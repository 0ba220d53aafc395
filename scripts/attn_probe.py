"""Attention-kernel probe: times drag_debug_attention on a packed batch of equal-length sequences."""
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
ap.add_argument("--variant", type=int, default=1)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
lib = _native.load()
heads, hd = 12, 32
T = a.seqs * a.len
cu = torch.arange(0, T + 1, a.len, dtype=torch.int32, device="cuda")
g = torch.Generator(device="cuda").manual_seed(5)
qkv = (torch.randn(T, 3 * heads * hd, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
ctx = torch.empty((T, heads * hd), device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream


def run():
    _native.check(lib.drag_debug_attention(0, a.variant, qkv.data_ptr(), ctx.data_ptr(), cu.data_ptr(), a.seqs, T, a.len, heads, st))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print(f"variant {a.variant}: {a.seqs} x {a.len}: {ms:.4f} ms, {4.0 * a.seqs * a.len * a.len * 384 / ms / 1e9:.0f} TFLOP/s, "
      f"{a.seqs * heads * a.len * a.len / ms / 1e6:.0f} Gexp/s")
