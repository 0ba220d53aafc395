#!/bin/bash
# round 2 GPU pass: attention v3 parity + probes, encoder parity incl. the outlier weights, model-dir path, retriever
mkdir -p gpurun_out
bash scripts/gpu_attn3.sh
echo "=== encoder (hf_init, stress, outlier)"
timeout 600 python -m pytest tests/test_encoder_gpu.py -q -k "not gemm and not attention" > gpurun_out/enc.log 2>&1; echo "exit=$?"; tail -n 12 gpurun_out/enc.log
echo "=== model dir + retriever"
timeout 600 python -m pytest tests/test_model_dir.py tests/test_retriever_gpu.py -q -m gpu > gpurun_out/retriever.log 2>&1; echo "exit=$?"; tail -n 12 gpurun_out/retriever.log
