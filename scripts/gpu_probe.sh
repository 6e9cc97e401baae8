#!/bin/bash
# Timing attribution of the scoring kernel: FM_PROBE=1 (max-tree fast path only), =2 (TMA+MMA pipeline only).
mkdir -p gpurun_out
for p in 0 1 2; do
  FM_PROBE=$p timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/probe_$p.json 2> gpurun_out/probe_$p.err
  python - <<PY
import json
j=json.load(open("gpurun_out/probe_$p.json"))
print("probe $p: score_ms", j["roofline"]["kernel_ms"], "phases", j["phases_ms"])
PY
done
