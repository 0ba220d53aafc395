#!/bin/bash
# tcgen05 attention (third design): parity of every variant, then probe timings against the mma.sync kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for v in ${TEST_VARIANTS:-3 4 5 6}; do
  timeout 300 python -m pytest tests/test_encoder_gpu.py -q -k "attention and ${v}-" > gpurun_out/attn3_v$v.log 2>&1
  echo "attention tests variant $v exit=$?"; grep -v "^drag_b200" gpurun_out/attn3_v$v.log | tail -n 4; grep "^drag_b200" gpurun_out/attn3_v$v.log | sort | uniq -c | head -5
done
for v in ${VARIANTS:-0 3 4 5 6}; do
  for shape in "1024 256" "2048 128" "512 512" "4096 64" "1 512"; do
    set -- $shape
    timeout 120 python scripts/attn_probe.py --variant $v --seqs $1 --len $2 --iters 20 > gpurun_out/probe.tmp 2>&1; rc=$?
    if [ $rc -ne 0 ]; then echo "variant $v: $1 x $2 FAILED rc=$rc: $(grep -v '^$' gpurun_out/probe.tmp | tail -n 2 | tr '\n' ' ')"; else tail -n 1 gpurun_out/probe.tmp; fi
  done
done
