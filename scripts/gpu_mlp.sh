#!/bin/bash
# fused feed-forward kernel (csrc/drag_mlp.cuh): parity, timing against the two GEMM kernels it replaces, in-kernel wait trace
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_encoder_gpu.py -q -x -k "fused_mlp" > gpurun_out/mlp_test.log 2>&1; echo "mlp tests exit=$?"; grep -v "^drag_b200" gpurun_out/mlp_test.log | tail -n 4; grep "^drag_b200" gpurun_out/mlp_test.log | sort | uniq -c | head -5
for t in 262144 65536 16384 12288 8192; do timeout 120 python scripts/mlp_probe.py --tokens $t 2>&1 | tail -n 2; done
# DRAG_MLP_DBG: 1 E1 without arithmetic, 3 also without the tensor-memory load, 4 G1 strictly two chunks ahead across tiles
for d in ${DBG:-}; do echo "DRAG_MLP_DBG=$d"; DRAG_MLP_DBG=$d timeout 120 python scripts/mlp_probe.py --tokens 262144 --what fused 2>&1 | tail -n 1; done
DRAG_MLP_TRACE=1 timeout 120 python scripts/mlp_probe.py --tokens 262144 --what fused --iters 2 2>&1 | grep "mlp trace" | tail -n 1
