#!/bin/bash
# Round 2: ncu --set full of one launch each of the compaction, rescoring and scoring kernels on the C4 workload
# (batches of 24 M rows), reports under gpurun_out/ (read back with scripts/ncu_summary.py).
mkdir -p gpurun_out
tag=${1:-r2}
for k in compact_kernel rescore_kernel score_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/${tag}_ncu_$k \
     python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline --no-wall > gpurun_out/${tag}_ncu_$k.log 2>&1
  tail -1 gpurun_out/${tag}_ncu_$k.log | cut -c1-160
done
ls -la gpurun_out/*.ncu-rep
