"""Attention-kernel debug: one launch of drag_debug_attention on given lengths, max error vs torch per sequence."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ai-dial-rag_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

from dial_rag_b200 import _native

variant = int(sys.argv[1])
lens = [int(x) for x in sys.argv[2].split(",")]
heads = int(sys.argv[3]) if len(sys.argv) > 3 else 12
hd = 32
lib = _native.load()
cu = np.zeros(len(lens) + 1, dtype=np.int32)
cu[1:] = np.cumsum(lens)
T = int(cu[-1])
g = torch.Generator(device="cuda").manual_seed(5)
qkv = (torch.randn(T, 3 * heads * hd, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
ctx = torch.full((T, heads * hd), float("nan"), device="cuda", dtype=torch.bfloat16)
d_cu = torch.from_numpy(cu).cuda()
import time
t_start = time.time()
_native.check(lib.drag_debug_attention(0, variant, qkv.data_ptr(), ctx.data_ptr(), d_cu.data_ptr(), len(lens), T, max(lens), heads,
                                       torch.cuda.current_stream().cuda_stream))
try:
    torch.cuda.synchronize()
except Exception as e:
    print(f"FAILED after {time.time() - t_start:.2f}s: {str(e).splitlines()[0]}")
    sys.exit(1)
print(f"ok in {time.time() - t_start:.3f}s")
got = ctx.float()
for i, n in enumerate(lens):
    blk = qkv[cu[i]:cu[i + 1]].float().view(n, 3, heads, hd)
    q, k, v = (blk[:, j].permute(1, 0, 2) for j in range(3))
    p = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(hd), dim=-1)
    ref = (p @ v).permute(1, 0, 2).reshape(n, heads * hd)
    d = (got[cu[i]:cu[i + 1]] - ref).abs()
    print(f"len {n}: max err {d.max().item():.4f} nan {int(torch.isnan(d).sum())} / {d.numel()}  per-head max {[round(x, 3) for x in d.view(n, heads, hd).amax((0, 2)).tolist()]}")
