#!/bin/bash
# bench.py on the big workloads (one GPU, final kernels)
mkdir -p gpurun_out
for w in c3 c4; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w rc=$?"; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/bench_$w.json"))
    print("$w value %.4e (%.2f ms) e2e %.4e (%.2f ms) score_ms %.3f frac %.3f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["kernel_ms"], j["roofline"]["frac"]), j["phases_ms"], j["rows_exact_frac"], j["compaction"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_$w.err").read()[-1500:])
PY
done
