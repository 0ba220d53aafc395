#!/bin/bash
mkdir -p gpurun_out
./scripts/ubench/softmax_mix 2>&1 | grep "scheduled loop" | tee gpurun_out/softmax_mix_scheduled.txt
timeout 300 python -m pytest tests/test_encoder_gpu.py -q -k "attention" > gpurun_out/attn_all.log 2>&1; echo "attention tests exit=$?"; grep -v "^drag_b200" gpurun_out/attn_all.log | tail -n 3
for shape in "1024 256" "2048 128" "512 512" "4096 64" "4096 40"; do set -- $shape; for v in 0 3; do timeout 120 python scripts/attn_probe.py --variant $v --seqs $1 --len $2 --iters 20 2>&1 | tail -n 1; done; done
DRAG_ATTN_STAGGER=0 timeout 120 python scripts/attn_trace.py --seqs 1024 --len 256 2>&1 | tail -n 5
DRAG_ATTN_STAGGER=1500 timeout 120 python scripts/attn_trace.py --seqs 1024 --len 256 2>&1 | tail -n 5
