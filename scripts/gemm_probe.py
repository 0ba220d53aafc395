"""GEMM-kernel probe: times drag_debug_gemm (variant 0/1/2, +10 = CTA pairs) on the encoder's shapes."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ai-dial-rag_b200")):
    sys.path.insert(0, p)
import torch

from dial_rag_b200 import _native

ap = argparse.ArgumentParser()
ap.add_argument("--variant", type=int, default=1)
ap.add_argument("--M", type=int, default=262144)
ap.add_argument("--N", type=int, default=1536)
ap.add_argument("--K", type=int, default=384)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
lib = _native.load()
M, N, K = a.M, a.N, a.K
g = torch.Generator(device="cuda").manual_seed(1)
dt = torch.float16 if a.variant >= 20 else torch.bfloat16   # +20: fp16 operands (FFN-down)
A = (torch.randn(M, K, device="cuda", generator=g) + 0.3).to(dt)
W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(dt)
colc = torch.randn(N, device="cuda", generator=g) * 0.2
cold = torch.randn(N, device="cuda", generator=g) * 0.1
gamma = torch.rand(N, device="cuda", generator=g) + 0.5
res = (torch.randn(M, N, device="cuda", generator=g)).to(torch.bfloat16) if a.variant % 10 == 2 else None
stats = torch.zeros((M, 3, 2), device="cuda")
stats[:, 0, 0] = 0.3 * (N if a.variant % 10 == 2 else K)
stats[:, 0, 1] = 1.09 * (N if a.variant % 10 == 2 else K)
out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
out_stats = torch.empty((M, 3, 2), device="cuda")
st = torch.cuda.current_stream().cuda_stream


def run():
    _native.check(lib.drag_debug_gemm(0, a.variant, A.data_ptr(), W.data_ptr(), colc.data_ptr(), cold.data_ptr(), gamma.data_ptr(),
                                      stats.data_ptr(), res.data_ptr() if res is not None else 0, out.data_ptr(), out_stats.data_ptr(),
                                      M, N, K, 1.0 / (N if a.variant % 10 == 2 else K), 1e-12, st))


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
print(f"variant {a.variant} dbg={os.environ.get('DRAG_GEMM_DBG', '0')}: M={M} N={N} K={K}: {ms:.4f} ms, {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s")
