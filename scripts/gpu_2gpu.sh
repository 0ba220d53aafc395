#!/bin/bash
# two GPUs of one box: the NCCL tests, then both bench arms under torchrun
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py -q 2>&1 | tail -n 4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 > gpurun_out/2gpu_reference.json 2> gpurun_out/2gpu_reference.err; echo "reference exit=$?"; head -c 400 gpurun_out/2gpu_reference.json; echo
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 > gpurun_out/2gpu_bench.json 2> gpurun_out/2gpu_bench.err; echo "bench exit=$?"; tail -n 3 gpurun_out/2gpu_bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/2gpu_bench.json") if l.startswith("{")][-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), "n_gpus", d["n_gpus"])
    print("parity", d["parity"])
    print({k: v for k, v in d["roofline"].items() if k.startswith("sharded")})
    print(d["extra"].get("search_sharded"))
except Exception as e:
    print("failed", e)
PY
