#!/bin/bash
# attention kernels: parity, then probe timings of every variant
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_encoder_gpu.py -q -x -k attention > gpurun_out/attn.log 2>&1
echo "attention tests exit=$?"; tail -n 15 gpurun_out/attn.log
for v in ${VARIANTS:-0}; do
  timeout 120 python scripts/attn_probe.py --variant $v --seqs 1024 --len 256 2>&1 | tail -n 2
  timeout 120 python scripts/attn_probe.py --variant $v --seqs 2048 --len 128 2>&1 | tail -n 1
  timeout 120 python scripts/attn_probe.py --variant $v --seqs 512 --len 512 2>&1 | tail -n 1
done
