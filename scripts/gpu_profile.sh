#!/bin/bash
# ncu evidence for the encoder step: plain run first (must exit 0), then the launch list and a
# --set full capture of the kernels named in $KERNELS (regex).  Never a bench value.
mkdir -p gpurun_out
TAG=${TAG:-prof}
CMD="python bench.py --steps 2 --warmup 3 --no-search --no-cpu-baseline --chunks ${CHUNKS:-256}"
timeout 600 $CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -n 20 gpurun_out/${TAG}_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "launch list exit=$?"
# skip the 3 warm-up forwards (62 launches each) and capture one layer's worth of each kernel
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${KERNELS:-attention_kernel|gemm_kernel}" -s ${SKIP:-150} -c ${COUNT:-5} \
    -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full exit=$?"; tail -n 3 gpurun_out/${TAG}_ncu_full.log
