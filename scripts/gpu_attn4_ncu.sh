#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/attn_probe.py --variant 4 --seqs 1024 --len 256 --iters 2 > gpurun_out/attn4_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc4 -s 3 -c 1 -f -o gpurun_out/attn4_full \
    python scripts/attn_probe.py --variant 4 --seqs 1024 --len 256 --iters 2 > gpurun_out/attn4_ncu.log 2>&1
echo "ncu exit=$?"; tail -n 2 gpurun_out/attn4_ncu.log
