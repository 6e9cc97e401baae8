#!/bin/bash
# ncu --set full of the scoring kernel (normal and FM_PROBE variants).  Usage: gpu_ncu_score.sh <tag> [probe...]
tag=${1:-x}; shift
mkdir -p gpurun_out
for p in ${@:-0}; do
  FM_PROBE=$p timeout 600 ncu --set full --clock-control none --import-source on -k regex:^score_kernel -s 2 -c 1 -f -o gpurun_out/prof_${tag}_p$p \
     python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${tag}_p$p.log 2>&1
  tail -2 gpurun_out/ncu_${tag}_p$p.log | cut -c1-200
done
