"""Scan vs batched path for small query counts over a 1M x 384 fp32 index (picks the crossover for batch_min_queries)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ai-dial-rag_b200")):
    sys.path.insert(0, p)
import torch

from dial_rag_b200.device_index import DeviceMatrix

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(2)
mat = torch.randn((rows, 384), generator=g, device=dev)
mat /= mat.norm(dim=1, keepdim=True)
dm = DeviceMatrix(mat)
for nq in (1, 2, 3, 4, 6, 8, 16):
    q = torch.randn((nq, 384), generator=g, device=dev)
    q = (q / q.norm(dim=1, keepdim=True)).double()
    res = {}
    for name, allow in (("scan", False), ("batch", True)):
        dm.batch_min_queries = 1 if allow else 10 ** 9
        for _ in range(3):
            dm.topk_device(q, 20, "sqeuclidean_dist")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dm.topk_device(q, 20, "sqeuclidean_dist")
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 10
    print(f"rows={rows} Q={nq}: scan {res['scan']:.3f} ms, batch {res['batch']:.3f} ms")
