#!/bin/bash
# Runs the GPU parity suites as separate processes (a faulting kernel poisons its CUDA context)
# and leaves one log per suite under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name"
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit=$? $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ')"
}
run search 900 python -m pytest tests/test_search_gpu.py -q -x
run gemm 300 python -m pytest tests/test_encoder_gpu.py -q -k gemm
run attention 200 python -m pytest tests/test_encoder_gpu.py -q -k attention
run encoder 600 python -m pytest tests/test_encoder_gpu.py -q -k "not gemm and not attention"
run retriever 300 python -m pytest tests/test_retriever_gpu.py -q
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
