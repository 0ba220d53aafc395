#!/bin/bash
# round 2 GPU pass e: attention v3 with P in tensor memory
mkdir -p gpurun_out
TEST_VARIANTS="3" VARIANTS="0 3" bash scripts/gpu_attn3.sh
bash scripts/gpu_trace.sh
echo "=== encoder suite with tc3"
DRAG_ATTENTION=tc3 timeout 600 python -m pytest tests/test_encoder_gpu.py -q -k "not gemm and not attention" > gpurun_out/enc_tc3.log 2>&1; echo "exit=$?"; tail -n 4 gpurun_out/enc_tc3.log
for v in tc3; do
  echo "=== bench, encoder only, DRAG_ATTENTION=$v"
  DRAG_ATTENTION=$v timeout 600 python bench.py --no-search --no-cpu-baseline --no-library-baseline > gpurun_out/bench_enc_$v.json 2> gpurun_out/bench_enc_$v.err; echo "exit=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_enc_$v.json"))
    print("$v", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), {k: round(x["avg_ms"], 4) for k, x in d["extra"]["kernels"].items()})
except Exception as e:
    print("$v", "failed", e)
PY
done
