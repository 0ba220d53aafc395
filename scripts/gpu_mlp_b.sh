#!/bin/bash
# fused feed-forward kernel inside the encoder: parity suites, A/B inside the step, ncu of the kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -q -k "not gemm_vs_torch and not attention_vs" > gpurun_out/mlp_enc.log 2>&1; echo "encoder suite exit=$?"; grep -v "^drag_b200" gpurun_out/mlp_enc.log | tail -n 8
for v in 0 1; do
  echo "=== bench, encoder only, DRAG_FUSED_MLP=$v"
  DRAG_FUSED_MLP=$v timeout 600 python bench.py --no-search --no-cpu-baseline --no-library-baseline > gpurun_out/bench_mlp_$v.json 2> gpurun_out/bench_mlp_$v.err; echo "exit=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_mlp_$v.json"))
    print("fused=$v", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), {k: round(x["avg_ms"], 4) for k, x in d["extra"]["kernels"].items()}, d["parity"], d["roofline"]["frac"])
except Exception as e:
    print("fused=$v", "failed", e)
PY
done
timeout 120 python scripts/mlp_probe.py --tokens 65536 --iters 3 --what fused > gpurun_out/mlp_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_kernel -s 3 -c 1 -f -o gpurun_out/mlp_full \
    python scripts/mlp_probe.py --tokens 65536 --iters 3 --what fused > gpurun_out/mlp_ncu.log 2>&1
echo "ncu exit=$?"; tail -n 2 gpurun_out/mlp_ncu.log
