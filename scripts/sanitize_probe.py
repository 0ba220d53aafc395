"""One small launch of every hand-rolled mbarrier / TMA / tcgen05 kernel, for compute-sanitizer:

    compute-sanitizer --tool racecheck python scripts/sanitize_probe.py

GEMM template (single CTA and CTA pair, the three epilogues), fused feed-forward kernel, both attention kernels, the
single-query ring scan and the batched tensor-core search (score / refine / rerank)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ai-dial-rag_b200")):
    sys.path.insert(0, p)
from dial_rag_b200 import _native  # noqa: E402
from dial_rag_b200.device_index import DeviceMatrix  # noqa: E402

lib = _native.load()
g = torch.Generator(device="cuda").manual_seed(0)
stream = torch.cuda.current_stream().cuda_stream
M, H, F = 300, 384, 1536


def stats_of(x):
    st = torch.zeros(x.shape[0], 3, 2, device="cuda")
    st[:, 0, 0] = x.float().sum(1)
    st[:, 0, 1] = (x.float() ** 2).sum(1)
    return st


x = (torch.randn(M, H, device="cuda", generator=g) + 0.3).to(torch.bfloat16)
st = stats_of(x)
for variant, N, K in ((0, 1152, 384), (10, 1152, 384), (1, 1536, 384), (11, 1536, 384), (2, 384, 384), (12, 384, 384), (32, 384, 1536)):
    f16 = variant >= 20
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.05
    a, w = (a.half(), w.half()) if f16 else (a.to(torch.bfloat16), w.to(torch.bfloat16))
    colc = torch.randn(N, device="cuda", generator=g)
    cold = torch.randn(N, device="cuda", generator=g)
    gamma = torch.rand(N, device="cuda", generator=g) + 0.5
    res = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
    normed = res if variant % 10 == 2 else a
    out = torch.empty(M, N, device="cuda", dtype=torch.float16 if variant % 10 == 1 else torch.bfloat16)
    ost = torch.zeros(M, 3, 2, device="cuda")
    _native.check(lib.drag_debug_gemm(0, variant, a.data_ptr(), w.data_ptr(), colc.data_ptr(), cold.data_ptr(), gamma.data_ptr(),
                                      stats_of(normed).data_ptr(), res.data_ptr(), out.data_ptr(), ost.data_ptr(), M, N, K,
                                      1.0 / normed.shape[1], 1e-12, stream))
    torch.cuda.synchronize()
    print("gemm variant", variant, "ok", flush=True)

w1 = (torch.randn(F, H, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
w2 = (torch.randn(H, F, device="cuda", generator=g) * 0.05).half()
up_c, up_d = w1.float().sum(1).contiguous(), torch.randn(F, device="cuda", generator=g)
cold, gamma = torch.randn(H, device="cuda", generator=g), torch.rand(H, device="cuda", generator=g) + 0.5
out = torch.empty(600, H, device="cuda", dtype=torch.bfloat16)
xx = (torch.randn(600, H, device="cuda", generator=g) + 0.3).to(torch.bfloat16)   # three 256-token tiles: two on one CTA pair at 1 pair... see grid
_native.check(lib.drag_debug_mlp(0, xx.data_ptr(), stats_of(xx).data_ptr(), w1.data_ptr(), up_c.data_ptr(), up_d.data_ptr(), w2.data_ptr(),
                                 cold.data_ptr(), gamma.data_ptr(), out.data_ptr(), torch.zeros(600, 3, 2, device="cuda").data_ptr(), 600, 1e-12, stream))
torch.cuda.synchronize()
print("fused mlp ok", flush=True)

lens = [300, 77, 512, 5]
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
T = int(sum(lens))
qkv = torch.randn(T, 3 * H, device="cuda", generator=g).to(torch.bfloat16)
ctx = torch.empty(T, H, device="cuda", dtype=torch.bfloat16)
for variant in (0, 3):
    _native.check(lib.drag_debug_attention(0, variant, qkv.data_ptr(), ctx.data_ptr(), cu.data_ptr(), len(lens), T, max(lens), 12, stream))
    torch.cuda.synchronize()
    print("attention variant", variant, "ok", flush=True)

rng = np.random.default_rng(0)
m = rng.standard_normal((20000, 384)).astype(np.float32)
dm = DeviceMatrix(m)
q = rng.standard_normal((1, 384))
dm.topk(q, 20, "inner_product")
print("single-query scan ok", flush=True)
dm.topk(rng.standard_normal((160, 384)), 10, "sqeuclidean_dist")
print("batched search ok", flush=True)
