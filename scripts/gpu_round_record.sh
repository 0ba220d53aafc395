#!/bin/bash
# Round-2 record: both bench arms on the final tree, then the ncu evidence (launch list + --set full of the GEMM, fused
# feed-forward and attention kernels) and a --set full capture of the tcgen05 attention kernel at the lengths it serves.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
echo "=== bench --impl reference"
timeout 900 python bench.py --impl reference > gpurun_out/k_bench_reference.json 2> gpurun_out/k_bench_reference.err; echo "exit=$?"
echo "=== bench (default)"
timeout 1500 python bench.py > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "exit=$?"; tail -n 3 gpurun_out/k_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/k_bench.json"))
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3), "parity", d["parity"]["all_ok"])
    print("kernels", {k: round(x["avg_ms"], 4) for k, x in d["extra"]["kernels"].items()})
except Exception as e:
    print("failed", e)
PY
echo "=== ncu"
TAG=k_prof KERNELS="attention_kernel|gemm_kernel|mlp_kernel" COUNT=5 SKIP=120 bash scripts/gpu_profile.sh
timeout 120 python scripts/attn_probe.py --variant 3 --seqs 148 --len 512 --iters 2 > gpurun_out/k_attn3_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc3 -s 3 -c 1 -f -o gpurun_out/k_attn3_full \
    python scripts/attn_probe.py --variant 3 --seqs 148 --len 512 --iters 2 > gpurun_out/k_attn3_ncu.log 2>&1
echo "ncu attn3 exit=$?"
timeout 120 python scripts/mlp_probe.py --tokens 262144 --iters 3 --what fused > gpurun_out/k_mlp_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_kernel -s 3 -c 1 -f -o gpurun_out/k_mlp_full \
    python scripts/mlp_probe.py --tokens 262144 --iters 3 --what fused > gpurun_out/k_mlp_ncu.log 2>&1
echo "ncu mlp exit=$?"
