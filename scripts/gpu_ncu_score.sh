#!/bin/bash
# ncu --set full of the scoring kernel on C2 (normal and probe builds).  Usage: gpu_ncu_score.sh <tag> [probe...]
tag=${1:-x}; shift
mkdir -p gpurun_out
for p in ${@:-0}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^score_kernel -s 2 -c 1 -f -o gpurun_out/prof_${tag}_p$p \
     python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-wall --debug-opt probe=$p > gpurun_out/ncu_${tag}_p$p.log 2>&1
  tail -2 gpurun_out/ncu_${tag}_p$p.log | cut -c1-200
done
