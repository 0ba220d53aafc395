#!/bin/bash
# pipe microbenchmarks + encoder kernel times vs batch size (does the activation working set fit L2?)
mkdir -p gpurun_out
timeout 120 scripts/ubench/pipes > gpurun_out/pipes.log 2>&1; echo "pipes exit=$?"; cat gpurun_out/pipes.log
for c in 32 64 128 256 1024; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-search --no-cpu-baseline --chunks $c > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  python - $c <<'PY'
import json, sys
c = int(sys.argv[1])
d = json.loads(open(f"gpurun_out/bench_c{c}.json").read())
k = d["extra"]["kernels"]
print("chunks", c, "chunks/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"],
      {n: round(v["avg_ms"] * 1024 / c, 4) for n, v in k.items()}, "(ms scaled to 1024 chunks)")
PY
done
