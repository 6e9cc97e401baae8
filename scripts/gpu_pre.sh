#!/bin/bash
# GPU parity tests, then bench with several look-ahead depths (FM_PRE)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for p in ${@:-0 4 8 12}; do
  FM_PRE=$p timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/pre_$p.json 2> gpurun_out/pre_$p.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/pre_$p.json"))
    print("FM_PRE=$p: value %.3e e2e %.3e score_ms %.4f frac %.3f scored %.3f" % (j["value"], j["e2e"]["value"], j["roofline"]["kernel_ms"], j["roofline"]["frac"], j["roofline"]["scored_fraction"]), {k: round(v,4) for k,v in j["phases_ms"].items()}, "exact_frac", j["rows_exact_frac"], "cand/row", j["candidates_per_row"])
except Exception as e:
    print("FM_PRE=$p failed", e); print(open("gpurun_out/pre_$p.err").read()[-2000:])
PY
done
