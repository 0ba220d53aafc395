#!/bin/bash
# final state of the tree: the whole GPU suite, smoke(), both bench arms
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "exit=$?"; grep -v "^drag_b200" gpurun_out/final_pytest_gpu.log | tail -n 4
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
echo "=== bench --impl reference"
timeout 900 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "exit=$?"
echo "=== bench (default)"
timeout 1700 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "exit=$?"; tail -n 3 gpurun_out/final_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/final_bench.json"))
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3), "parity", d["parity"])
    print("kernels", {k: round(x["avg_ms"], 4) for k, x in d["extra"]["kernels"].items()})
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k not in ("api", "search_note")})
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"].items() if k.startswith(("search", "query"))})
    print(d["extra"].get("search_error"), d["clocks"])
except Exception as e:
    print("failed", e)
PY
