#!/bin/bash
timeout 300 python -m pytest tests/test_encoder_gpu.py -q -k "attention" 2>&1 | grep -v "^drag_b200" | tail -n 2
for ns in 0 1500 2500 4000 8000; do echo "stagger $ns ns"; for shape in "1024 256" "2048 128"; do set -- $shape; DRAG_ATTN4_STAGGER_NS=$ns timeout 120 python scripts/attn_probe.py --variant 4 --seqs $1 --len $2 --iters 20 2>&1 | tail -n 1; done; done
