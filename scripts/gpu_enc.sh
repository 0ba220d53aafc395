#!/bin/bash
# encoder parity (kernel-level + end-to-end) and the encoder-only bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_retriever_gpu.py -q -x > gpurun_out/enc.log 2>&1
echo "encoder tests exit=$?"; tail -n 4 gpurun_out/enc.log
timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 --no-search --no-cpu-baseline > gpurun_out/bench_enc.json 2> gpurun_out/bench_enc.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_enc.json").read())
print("chunks/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"],
      {k: (round(v["avg_ms"], 4), round(v.get("tflops", 0))) for k, v in d["extra"]["kernels"].items()})
PY
