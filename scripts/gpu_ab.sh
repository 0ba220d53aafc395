#!/bin/bash
# same-box A/B of encoder configurations: DRAG_GEMM_PAIRS masks given as arguments, alternated ROUNDS times
mkdir -p gpurun_out
for r in $(seq 1 ${ROUNDS:-2}); do
  for m in "$@"; do
    DRAG_GEMM_PAIRS=$m timeout 300 python bench.py --steps 5 --warmup 3 --no-search --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('mask $m chunks/s', round(d['value']), {k:(round(v['avg_ms'],4)) for k,v in d['extra']['kernels'].items() if k.startswith('gemm') or k=='attention'})"
  done
done
