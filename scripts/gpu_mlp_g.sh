#!/bin/bash
# whole tree after the fused feed-forward kernel: GPU test suite, threshold probe, default bench
mkdir -p gpurun_out
for t in 12288; do timeout 120 python scripts/mlp_probe.py --tokens $t --iters 50 2>&1 | tail -n 2; done
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest_gpu.log 2>&1; echo "exit=$?"; grep -v "^drag_b200" gpurun_out/g_pytest_gpu.log | tail -n 8
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
echo "=== bench (default)"
timeout 1500 python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "exit=$?"; tail -n 5 gpurun_out/g_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/g_bench.json"))
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3))
    print("roofline frac", d["roofline"]["frac"], "achieved", d["roofline"]["achieved"]); print("parity", d.get("parity"))
    print("kernels", {k: round(x["avg_ms"], 4) for k, x in d["extra"]["kernels"].items()})
    print("e2e", d["e2e"])
except Exception as e:
    print("failed", e)
PY
