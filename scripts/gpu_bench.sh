#!/bin/bash
# bench.py (plain), then the ncu launch list of the same command (shares only, never a bench value)
mkdir -p gpurun_out
STEPS=${STEPS:-5}; WARM=${WARM:-3}
timeout 900 python bench.py --steps $STEPS --warmup $WARM > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?"; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ "${NCU:-0}" = "1" ]; then
  timeout 600 python bench.py --steps 2 --warmup 1 --no-search --no-cpu-baseline --chunks 256 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-search --no-cpu-baseline --chunks 256 > gpurun_out/ncu_launch.log 2>&1
  echo "ncu exit=$?"
fi
