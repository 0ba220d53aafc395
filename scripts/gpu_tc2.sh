#!/bin/bash
# tc2 attention: per-length errors, parity test, probe timings
mkdir -p gpurun_out
for l in 64 256 1,2,17,64,65,128,256,300,512,33; do
  echo "== lens $l"; timeout 120 python scripts/attn_debug.py 2 $l 2>&1 | grep -E "FAILED|ok in|max err|drag_b200" | head -12
done
VARIANTS="2 0" bash scripts/gpu_attn.sh
