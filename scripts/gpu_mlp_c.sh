#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./scripts/ubench/umma_rate 2>&1 | tee gpurun_out/umma_rate.txt
for v in mma tc3; do
  echo "=== layer taps, outlier weights, DRAG_ATTENTION=$v"
  DRAG_ATTENTION=$v timeout 300 python -m pytest tests/test_encoder_gpu.py -q -k "layer_taps" 2>&1 | grep -E "AssertionError: \(|passed|failed" | head -5
done
