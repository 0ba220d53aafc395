"""One 512-token query through the encoder (device resident): wall/device time per forward; run under ncu for the launch list."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ai-dial-rag_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

from dial_rag_b200.embeddings.encoder import B200Encoder
import synth_weights

n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 1
q_len = int(sys.argv[2]) if len(sys.argv) > 2 else 512
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dev = torch.device("cuda", 0)
enc = B200Encoder(synth_weights.synth_weights(seed=0, style="hf_init"), device=0, max_tokens=n_seq * q_len)
rng = np.random.default_rng(5)
ids = rng.integers(1000, 30522, size=(n_seq, q_len), dtype=np.int32)
ids[:, 0], ids[:, -1] = 101, 102
cu = np.arange(0, n_seq * q_len + 1, q_len, dtype=np.int32)
d_ids, d_cu = torch.from_numpy(ids.reshape(-1)).to(dev), torch.from_numpy(cu).to(dev)
d_out = torch.empty((n_seq, 384), dtype=torch.float32, device=dev)
for _ in range(3):
    enc.forward_device(d_ids, d_cu, cu, d_out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(iters):
    enc.forward_device(d_ids, d_cu, cu, d_out)
e1.record()
torch.cuda.synchronize()
print(f"{n_seq} x {q_len} tokens: {e0.elapsed_time(e1) / iters:.4f} ms per forward on the device, "
      f"{1e3 * (time.perf_counter() - t0) / iters:.4f} ms wall")
