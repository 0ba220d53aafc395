"""Batched-search probe: times drag_topk_batch on a synthetic on-device index (for ncu launch lists)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ai-dial-rag_b200")):
    sys.path.insert(0, p)
import torch

from dial_rag_b200.device_index import DeviceMatrix

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--queries", type=int, default=1000)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--metric", default="inner_product")
ap.add_argument("--storage", default="f32")
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(2)
mat = torch.empty((a.rows, 384), dtype=torch.float32, device=dev)
for s in range(0, a.rows, 1 << 20):
    blk = torch.randn((min(1 << 20, a.rows - s), 384), generator=g, device=dev)
    mat[s:s + (1 << 20)] = blk / blk.norm(dim=1, keepdim=True)
dm = DeviceMatrix(mat, storage=a.storage)
del mat
q = torch.randn((a.queries, 384), generator=g, device=dev)
q = (q / q.norm(dim=1, keepdim=True)).double()
dm.topk_device(q, a.k, a.metric)
torch.cuda.synchronize()
for _ in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dm.topk_device(q, a.k, a.metric)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rows={a.rows} Q={a.queries} k={a.k} {a.metric} {a.storage}: {ms:.3f} ms/batch, {a.queries / ms * 1e3:.0f} q/s, "
          f"{2.0 * a.queries * a.rows * 384 / ms / 1e9:.0f} TFLOP/s, fallbacks={dm.last_batch_fallbacks}", flush=True)
