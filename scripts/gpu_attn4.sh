#!/bin/bash
# one-job-per-CTA tcgen05 attention kernel (variant 4): parity, probe timings against variants 0 and 3, inside the step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_encoder_gpu.py -q -k "attention" > gpurun_out/attn4_test.log 2>&1; echo "attention tests exit=$?"; grep -v "^drag_b200" gpurun_out/attn4_test.log | tail -n 6; grep "^drag_b200" gpurun_out/attn4_test.log | sort | uniq -c | head -5
for v in 0 3 4; do for shape in "1024 256" "2048 128" "4096 64" "1 256"; do set -- $shape; timeout 120 python scripts/attn_probe.py --variant $v --seqs $1 --len $2 --iters 20 > gpurun_out/probe.tmp 2>&1; rc=$?; if [ $rc -ne 0 ]; then echo "variant $v: $1 x $2 FAILED rc=$rc: $(grep -v '^$' gpurun_out/probe.tmp | tail -n 2 | tr '\n' ' ')"; else tail -n 1 gpurun_out/probe.tmp; fi; done; done
for v in auto_mma tc4; do
  echo "=== bench, encoder only, DRAG_ATTENTION=$v"
  DRAG_ATTENTION=$v timeout 600 python bench.py --no-search --no-cpu-baseline --no-library-baseline > gpurun_out/bench_attn_$v.json 2> gpurun_out/bench_attn_$v.err; echo "exit=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_attn_$v.json"))
    print("$v", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), {k: round(x["avg_ms"], 4) for k, x in d["extra"]["kernels"].items()}, d["parity"])
except Exception as e:
    print("$v", "failed", e)
PY
done
