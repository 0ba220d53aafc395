#!/bin/bash
# round 2 GPU pass b: attention v3 (uniform issuers) parity + probes, GEMM / search-batch parity after the elect_one change,
# encoder + retriever suites, then the bench with both attention kernels
mkdir -p gpurun_out
bash scripts/gpu_attn3.sh
echo "=== gemm"
timeout 300 python -m pytest tests/test_encoder_gpu.py -q -k gemm > gpurun_out/gemm.log 2>&1; echo "exit=$?"; tail -n 3 gpurun_out/gemm.log
echo "=== search batch"
timeout 600 python -m pytest tests/test_search_batch_gpu.py -q > gpurun_out/search_batch.log 2>&1; echo "exit=$?"; tail -n 3 gpurun_out/search_batch.log
echo "=== encoder (hf_init, stress, outlier)"
timeout 600 python -m pytest tests/test_encoder_gpu.py -q -k "not gemm and not attention" > gpurun_out/enc.log 2>&1; echo "exit=$?"; tail -n 5 gpurun_out/enc.log
echo "=== model dir + retriever"
timeout 600 python -m pytest tests/test_model_dir.py tests/test_retriever_gpu.py -q -m gpu > gpurun_out/retriever.log 2>&1; echo "exit=$?"; tail -n 5 gpurun_out/retriever.log
echo "=== bench, encoder only, mma.sync attention"
timeout 600 python bench.py --no-search --no-cpu-baseline --no-library-baseline > gpurun_out/bench_enc_v0.json 2> gpurun_out/bench_enc_v0.err; echo "exit=$?"
python - <<'PY'
import json
for v in ("v0",):
    try:
        d = json.load(open(f"gpurun_out/bench_enc_{v}.json"))
        print(v, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), {k: round(x["avg_ms"], 4) for k, x in d["extra"]["kernels"].items()})
    except Exception as e:
        print(v, "failed", e)
PY
for v in tc3 tc3p2; do
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
