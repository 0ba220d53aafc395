#!/bin/bash
# float64 scan (drag_topk) bandwidth: rows x queries sweep, then one ncu --set full capture
mkdir -p gpurun_out
for spec in "1000000 1 20" "1000000 4 20" "10000000 1 100" "10000000 4 100"; do
  set -- $spec
  timeout 300 python scripts/search_probe.py --rows $1 --queries $2 --k $3 --iters 3 2>&1 | tail -n 1 | awk -v r=$1 '{ms=$6; printf "%s => %.0f GB/s\n", $0, r*384*4/ms/1e6}'
done
if [ "${NCU:-0}" = "1" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_vec -s 1 -c 1 -f -o gpurun_out/scan \
  python scripts/search_probe.py --rows 10000000 --queries 1 --k 100 --iters 1 > gpurun_out/scan_ncu.log 2>&1
echo "ncu exit=$?"
fi
