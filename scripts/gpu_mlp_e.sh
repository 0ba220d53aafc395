#!/bin/bash
mkdir -p gpurun_out
for d in 0 3; do echo "DRAG_MLP_DBG=$d"; DRAG_MLP_TRACE=1 DRAG_MLP_DBG=$d timeout 120 python scripts/mlp_probe.py --tokens 262144 --what fused --iters 2 2>&1 | tail -n 2; done
