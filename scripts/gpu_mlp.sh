#!/bin/bash
# fused feed-forward kernel: parity, then timing against the two GEMM kernels it replaces
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_encoder_gpu.py -q -x -k "fused_mlp" > gpurun_out/mlp_test.log 2>&1; echo "mlp tests exit=$?"; grep -v "^drag_b200" gpurun_out/mlp_test.log | tail -n 12; grep "^drag_b200" gpurun_out/mlp_test.log | sort | uniq -c | head -5
for t in 262144 65536 16384; do timeout 120 python scripts/mlp_probe.py --tokens $t 2>&1 | tail -n 2; done
echo "=== encoder: batch invariance with the per-sequence attention dispatch"
timeout 300 python -m pytest tests/test_encoder_gpu.py -q -k "batch_composition or peaked" 2>&1 | tail -n 3
