#!/bin/bash
# round 2 GPU pass f: attention v3 after the garbage-row fix; encoder suite with it; bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_encoder_gpu.py -q -k "attention" > gpurun_out/attn_all.log 2>&1; echo "attention tests exit=$?"; grep -v "^drag_b200" gpurun_out/attn_all.log | tail -n 6
for shape in "1024 256" "2048 128" "512 512"; do set -- $shape; timeout 120 python scripts/attn_probe.py --variant 3 --seqs $1 --len $2 --iters 20 2>&1 | tail -n 1; done
echo "=== encoder suite with tc3"
DRAG_ATTENTION=tc3 timeout 600 python -m pytest tests/test_encoder_gpu.py -q -k "not gemm and not attention" > gpurun_out/enc_tc3.log 2>&1; echo "exit=$?"; tail -n 5 gpurun_out/enc_tc3.log
echo "=== retriever + model dir with tc3"
DRAG_ATTENTION=tc3 timeout 600 python -m pytest tests/test_model_dir.py tests/test_retriever_gpu.py -q -m gpu > gpurun_out/retriever_tc3.log 2>&1; echo "exit=$?"; tail -n 4 gpurun_out/retriever_tc3.log
echo "=== scale + threads"
timeout 900 python -m pytest tests/test_scale_and_threads_gpu.py -q > gpurun_out/scale.log 2>&1; echo "exit=$?"; tail -n 6 gpurun_out/scale.log
