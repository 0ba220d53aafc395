#!/bin/bash
# GEMM parity tests + probe timings of the four encoder GEMMs (pair and single forms)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -q -x -k gemm > gpurun_out/gemm.log 2>&1; echo "gemm tests exit=$? $(tail -n 1 gpurun_out/gemm.log)"
run() { timeout 120 python scripts/gemm_probe.py --variant $1 --N $2 --K $3 2>&1 | tail -n 1; }
for v in ${VARIANTS:-10 11 32 2}; do
  case $v in
    0|10) run $v 1152 384;; 1|11) run $v 1536 384;; 2|12) run $v 384 384;; 22|32) run $v 384 1536;;
  esac
done
