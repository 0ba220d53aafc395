#!/bin/bash
# launch list (ncu gpu__time_duration) of one single-query top-20 over 1M rows and of one 512-token forward
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/scan_launches.csv \
  python scripts/search_probe.py --rows 1000000 --queries 1 --k 20 --iters 1 > gpurun_out/scan_launches.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/scan_launches.csv")) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
out = [(r[ki][:70], float(r[vi].replace(",", ""))) for r in rows[1:]]
# the last topk call = everything after the last query_prep kernel
start = max(i for i, (n, _) in enumerate(out) if "query_prep" in n)
for n, v in out[start:]:
    print(f"{n:72s} {v / 1000:8.1f} us")
PY
