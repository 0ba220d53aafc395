#!/bin/bash
# batched search: parity suite, then the search legs of bench.py
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_search_batch_gpu.py -x -q -s > gpurun_out/batch.log 2>&1
echo "batch tests exit=$?"; tail -n 25 gpurun_out/batch.log
