#!/bin/bash
# tcgen05 attention (third design): parity of every variant, then probe timings against the mma.sync kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -m pytest tests/test_encoder_gpu.py -q -x -k "attention and (3 or 4)" > gpurun_out/attn3.log 2>&1
echo "attention tests exit=$?"; tail -n 15 gpurun_out/attn3.log
for v in ${VARIANTS:-0 3 4}; do
  timeout 120 python scripts/attn_probe.py --variant $v --seqs 1024 --len 256 2>&1 | tail -n 2
  timeout 120 python scripts/attn_probe.py --variant $v --seqs 2048 --len 128 2>&1 | tail -n 1
  timeout 120 python scripts/attn_probe.py --variant $v --seqs 512 --len 512 2>&1 | tail -n 1
  timeout 120 python scripts/attn_probe.py --variant $v --seqs 1 --len 512 --iters 50 2>&1 | tail -n 1
done
