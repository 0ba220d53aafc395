#!/bin/bash
# pipe-rate microbenchmark of the softmax loop + ncu --set full (source-level stalls) of the tcgen05 attention kernel
mkdir -p gpurun_out
./scripts/ubench/softmax_mix > gpurun_out/softmax_mix.txt 2>&1; echo "softmax_mix exit=$?"; cat gpurun_out/softmax_mix.txt
V=${V:-3}
timeout 120 python scripts/attn_probe.py --variant $V --seqs 296 --len 256 --iters 2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_tc3 -s 3 -c 1 -o gpurun_out/attn3_prof \
    python scripts/attn_probe.py --variant $V --seqs 296 --len 256 --iters 2 > gpurun_out/ncu.log 2>&1
echo "ncu exit=$?"; tail -n 3 gpurun_out/ncu.log
