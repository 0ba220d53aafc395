#!/bin/bash
# one ncu --set full capture of the attention kernel on the probe workload (never a timing source)
mkdir -p gpurun_out
V=${VARIANT:-0}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention -s 3 -c 1 -f -o gpurun_out/attn_v$V \
  python scripts/attn_probe.py --variant $V --seqs ${SEQS:-1024} --len ${LEN:-256} --iters 2 > gpurun_out/attn_v${V}_ncu.log 2>&1
echo "ncu exit=$?"; tail -n 2 gpurun_out/attn_v${V}_ncu.log
