#!/bin/bash
for rep in 1 2; do for d in 0 4; do echo "DRAG_MLP_DBG=$d"; DRAG_MLP_DBG=$d timeout 120 python scripts/mlp_probe.py --tokens 262144 --what fused --iters 40 2>&1 | tail -n 1; done; done
