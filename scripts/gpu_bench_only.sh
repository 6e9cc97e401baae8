#!/bin/bash
# one bench line, no tests, no CPU baseline
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/bench_quick.json"))
print("value %.4e (%.3f ms)  e2e %.4e (%.3f ms)  score_ms %.4f frac %.3f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["kernel_ms"], j["roofline"]["frac"]), j["phases_ms"])
PY
