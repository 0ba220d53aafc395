#!/bin/bash
# compute-sanitizer racecheck over one small launch of every hand-rolled mbarrier/TMA/tcgen05 kernel (plain run first)
mkdir -p gpurun_out
timeout 300 python scripts/sanitize_probe.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -n 20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 40 python scripts/sanitize_probe.py > gpurun_out/racecheck.log 2>&1
echo "racecheck exit=$?"; grep -c "Race reported\|Error:" gpurun_out/racecheck.log; tail -n 25 gpurun_out/racecheck.log
