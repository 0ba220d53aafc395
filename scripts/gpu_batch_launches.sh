#!/bin/bash
# launch list (ncu gpu__time_duration) of one batched top-20 call (Q queries over 1M rows)
mkdir -p gpurun_out
Q=${Q:-8}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/batch_launches.csv \
  python scripts/search_probe.py --rows 1000000 --queries $Q --k 20 --iters 1 > gpurun_out/batch_launches.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/batch_launches.csv")) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
out = [(r[ki][:80], float(r[vi].replace(",", ""))) for r in rows[1:]]
# the last topk call = everything from the last query-side preparation kernel on
names = [n for n, _ in out]
start = max(i for i, n in enumerate(names) if "rerank" in n)
# walk back to the first score kernel of that call
first = start
while first > 0 and ("score" in names[first - 1] or "refine" in names[first - 1] or "prep" in names[first - 1] or "elementwise" in names[first - 1] or "query" in names[first - 1]):
    first -= 1
tot = 0
for n, v in out[first:start + 3]:
    print(f"{n:82s} {v / 1000:8.1f} us"); tot += v
print("sum", tot / 1000, "us")
PY
