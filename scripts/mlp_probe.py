"""Time the fused feed-forward kernel (drag_debug_mlp) against the two GEMM kernels it replaces (drag_debug_gemm 11 + 32).

    python scripts/mlp_probe.py [--tokens 262144] [--iters 20]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ai-dial-rag_b200"))
from dial_rag_b200 import _native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tokens", type=int, default=262144)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--what", default="both")
args = ap.parse_args()
lib = _native.load()
M, H, F = args.tokens, 384, 1536
g = torch.Generator(device="cuda").manual_seed(1)
# two activation buffers (rotated) so that consecutive launches do not hit the same lines in L2
xs = [(torch.randn(M, H, device="cuda", generator=g) + 0.3).to(torch.bfloat16) for _ in range(2)]
w1g = (torch.randn(F, H, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
up_c = w1g.float().sum(1).contiguous()
up_d = torch.randn(F, device="cuda", generator=g) * 0.3
w2 = (torch.randn(H, F, device="cuda", generator=g) * 0.05).half()
cold = torch.randn(H, device="cuda", generator=g) * 0.1
gamma = torch.rand(H, device="cuda", generator=g) + 0.5
stats = []
for x in xs:
    st = torch.zeros(M, 3, 2, device="cuda")
    st[:, 0, 0] = x.float().sum(1)
    st[:, 0, 1] = (x.float() ** 2).sum(1)
    stats.append(st)
out = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
out_stats = torch.zeros(M, 3, 2, device="cuda")
hbuf = torch.empty(M, F, device="cuda", dtype=torch.float16)
stream = torch.cuda.current_stream().cuda_stream
flop = 2.0 * 2 * M * H * F


def fused(i):
    x, st = xs[i & 1], stats[i & 1]
    _native.check(lib.drag_debug_mlp(0, x.data_ptr(), st.data_ptr(), w1g.data_ptr(), up_c.data_ptr(), up_d.data_ptr(), w2.data_ptr(),
                                     cold.data_ptr(), gamma.data_ptr(), out.data_ptr(), out_stats.data_ptr(), M, 1e-12, stream))


def unfused(i):
    x, st = xs[i & 1], stats[i & 1]
    _native.check(lib.drag_debug_gemm(0, 11, x.data_ptr(), w1g.data_ptr(), up_c.data_ptr(), up_d.data_ptr(), None, st.data_ptr(), None,
                                      hbuf.data_ptr(), None, M, F, H, 1.0 / H, 1e-12, stream))
    _native.check(lib.drag_debug_gemm(0, 32, hbuf.data_ptr(), w2.data_ptr(), None, cold.data_ptr(), gamma.data_ptr(), st.data_ptr(),
                                      x.data_ptr(), out.data_ptr(), out_stats.data_ptr(), M, H, F, 1.0 / H, 1e-12, stream))


for name, fn in (("fused", fused), ("unfused", unfused)):
    if args.what not in ("both", name):
        continue
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"{name}: {M} tokens: {ms:.4f} ms, {flop / ms / 1e9:.0f} TFLOP/s")
