#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_encoder_gpu.py -q -k "attention" > gpurun_out/attn_all.log 2>&1; echo "attention tests exit=$?"; grep -v "^drag_b200" gpurun_out/attn_all.log | tail -n 6; grep "^drag_b200" gpurun_out/attn_all.log | sort | uniq -c | head -5
for shape in "1024 256" "2048 128" "512 512" "4096 64" "1 512"; do set -- $shape; timeout 120 python scripts/attn_probe.py --variant 3 --seqs $1 --len $2 --iters 20 > gpurun_out/probe.tmp 2>&1; rc=$?; if [ $rc -ne 0 ]; then echo "variant 3: $1 x $2 FAILED rc=$rc: $(grep -v '^$' gpurun_out/probe.tmp | tail -n 2 | tr '\n' ' ')"; else tail -n 1 gpurun_out/probe.tmp; fi; done
timeout 120 python scripts/attn_trace.py --seqs 1024 --len 256 2>&1 | tail -n 22
echo "=== encoder suite with tc3"
DRAG_ATTENTION=tc3 timeout 600 python -m pytest tests/test_encoder_gpu.py -q -k "not gemm and not attention" > gpurun_out/enc_tc3.log 2>&1; echo "exit=$?"; tail -n 4 gpurun_out/enc_tc3.log
for v in mma tc3; do
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
