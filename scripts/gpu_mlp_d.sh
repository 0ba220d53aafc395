#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 3; do echo "DRAG_MLP_DBG=$d"; DRAG_MLP_DBG=$d timeout 120 python scripts/mlp_probe.py --tokens 262144 --what fused 2>&1 | tail -n 1; done
timeout 200 python -m pytest tests/test_encoder_gpu.py -q -x -k "fused_mlp or gemm_vs_torch" 2>&1 | tail -n 3
