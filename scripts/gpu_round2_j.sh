#!/bin/bash
# Validation of the whole tree: GPU test suite, smoke(), default bench (both arms), then ncu evidence.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest_gpu.log 2>&1; echo "exit=$?"; grep -v "^drag_b200" gpurun_out/j_pytest_gpu.log | tail -n 8
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 3
echo "=== bench --impl reference"
timeout 900 python bench.py --impl reference > gpurun_out/j_bench_reference.json 2> gpurun_out/j_bench_reference.err; echo "exit=$?"; head -c 600 gpurun_out/j_bench_reference.json; echo
echo "=== bench (default)"
timeout 1500 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "exit=$?"; tail -n 5 gpurun_out/j_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/j_bench.json"))
    print("value", round(d["value"]), "e2e", d["e2e"], "ms/step", round(d["ms_per_step"], 3))
    print("roofline", d["roofline"]); print("cpu_baseline", d["cpu_baseline"]); print("parity", d.get("parity"))
    print("kernels", {k: round(x["avg_ms"], 4) for k, x in d["extra"]["kernels"].items()})
    print("library", d.get("gpu_library_baseline"))
except Exception as e:
    print("failed", e)
PY
echo "=== ncu"
TAG=j_prof KERNELS="attention_kernel|attn3|gemm_kernel" COUNT=6 bash scripts/gpu_profile.sh
echo "=== ncu tcgen05 attention (512 x 512: the length range it is dispatched for)"
timeout 120 python scripts/attn_probe.py --variant 3 --seqs 148 --len 512 --iters 2 > gpurun_out/j_attn3_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc3 -s 3 -c 1 -f -o gpurun_out/j_attn3_full \
    python scripts/attn_probe.py --variant 3 --seqs 148 --len 512 --iters 2 > gpurun_out/j_attn3_ncu.log 2>&1
echo "ncu exit=$?"; tail -n 2 gpurun_out/j_attn3_ncu.log
