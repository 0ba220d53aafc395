#!/bin/bash
# probe timings of the four encoder GEMMs, then one ncu --set full capture each (never a timing source)
mkdir -p gpurun_out
run() { timeout 120 python scripts/gemm_probe.py --variant $1 --N $2 --K $3 2>&1 | tail -n 1; }
run 10 1152 384; run 2 384 384; run 11 1536 384; run 32 384 1536
for spec in "10 1152 384 qkv" "11 1536 384 up" "32 384 1536 down" "2 384 384 out"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/gemm_$4 \
    python scripts/gemm_probe.py --variant $1 --N $2 --K $3 --iters 2 > gpurun_out/gemm_$4_ncu.log 2>&1
  echo "ncu $4 exit=$?"
done
