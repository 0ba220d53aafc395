#!/bin/bash
# programmatic dependent launch of the forward's kernels: parity suites, then same-box A/B (DRAG_PDL=0 / 1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_retriever_gpu.py tests/test_scale_and_threads_gpu.py -q -x 2>&1 | grep -v "^drag_b200" | tail -n 4
for rep in 1 2; do for v in 0 1; do
  DRAG_PDL=$v timeout 600 python bench.py --no-cpu-baseline --no-library-baseline --search-rows 1000000 > gpurun_out/bench_pdl_$v.json 2> gpurun_out/bench_pdl_$v.err; echo "DRAG_PDL=$v exit=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_pdl_$v.json"))
    print("pdl=$v", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), "query batch1 device p50", round(d["roofline"].get("query_batch1_device_ms_p50", 0), 4),
          "host p50", round(d["e2e"].get("query_batch1_host_api_ms_p50", 0), 4), "batch256 qps", round(d["roofline"].get("query_batch256_qps_device", 0)), d["parity"]["all_ok"], d["extra"].get("search_error"))
except Exception as e:
    print("pdl=$v", "failed", e)
PY
done; done
