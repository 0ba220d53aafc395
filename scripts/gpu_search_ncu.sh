#!/bin/bash
# ncu --set full of the single-query ring scan and of the batched score kernel; only the text summaries travel back
mkdir -p gpurun_out /tmp/ncu
timeout 300 ncu --set full --clock-control none -k regex:scan_ring -s 1 -c 1 -f -o /tmp/ncu/scan_ring \
  python scripts/search_probe.py --rows 10000000 --queries 1 --k 100 --iters 1 > gpurun_out/scan_ring_ncu.log 2>&1; echo "ncu scan exit=$?"
timeout 300 ncu --set full --clock-control none -k regex:score_kernel -s 6 -c 1 -f -o /tmp/ncu/score \
  python scripts/search_probe.py --rows 10000000 --queries 1000 --k 100 --iters 1 > gpurun_out/score_ncu.log 2>&1; echo "ncu score exit=$?"
python scripts/ncu_summary.py /tmp/ncu/scan_ring.ncu-rep > gpurun_out/scan_ring_summary.txt 2>&1
python scripts/ncu_summary.py /tmp/ncu/score.ncu-rep > gpurun_out/score_summary.txt 2>&1
cat gpurun_out/scan_ring_summary.txt gpurun_out/score_summary.txt
