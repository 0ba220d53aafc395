#!/bin/bash
# round 2 GPU pass g: stagger sweep of the tcgen05 attention; full bench with it
mkdir -p gpurun_out
for sg in 0 600 1200 1800 2400; do
  for shape in "1024 256" "512 512" "2048 128"; do set -- $shape
    echo -n "stagger $sg: "; DRAG_ATTN_STAGGER=$sg timeout 120 python scripts/attn_probe.py --variant 3 --seqs $1 --len $2 --iters 20 2>&1 | tail -n 1
  done
done
DRAG_ATTN_STAGGER=1200 timeout 120 python scripts/attn_trace.py --seqs 1024 --len 256 2>&1 | tail -n 13
echo "=== full bench, DRAG_ATTENTION=tc3"
DRAG_ATTENTION=tc3 DRAG_ATTN_STAGGER=1200 timeout 900 python bench.py > gpurun_out/bench_full_tc3.json 2> gpurun_out/bench_full_tc3.err; echo "exit=$?"; tail -n 5 gpurun_out/bench_full_tc3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_full_tc3.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3))
print("kernels", {k: round(x["avg_ms"], 4) for k, x in d["extra"]["kernels"].items()})
print("parity", d.get("parity"))
print("cpu_baseline", d.get("cpu_baseline"))
print("gpu_library_baseline", d.get("gpu_library_baseline"))
print("roofline scalars", {k: v for k, v in d["roofline"].items() if not isinstance(v, (dict, str))})
print("e2e", d["e2e"])
print("extra keys", list(d["extra"].keys()), d["extra"].get("search_error"))
PY
