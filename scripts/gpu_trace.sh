#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/attn_trace.py --seqs 1024 --len 256 2>&1 | tail -n 16
timeout 120 python scripts/attn_trace.py --seqs 512 --len 512 2>&1 | tail -n 16
