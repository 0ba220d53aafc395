#!/bin/bash
# round 2, first GPU pass: attention v3 parity + probes, encoder parity incl. the outlier weights, model-dir path
mkdir -p gpurun_out
bash scripts/gpu_attn3.sh
echo "=== encoder (hf_init, stress, outlier)"
timeout 600 python -m pytest tests/test_encoder_gpu.py -q -k "not gemm and not attention" > gpurun_out/enc.log 2>&1; echo "exit=$?"; tail -n 25 gpurun_out/enc.log
echo "=== model dir"
timeout 300 python -m pytest tests/test_model_dir.py -q -m gpu > gpurun_out/model_dir.log 2>&1; echo "exit=$?"; tail -n 15 gpurun_out/model_dir.log
