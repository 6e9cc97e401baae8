#!/bin/bash
# Compaction version 2 (pipelined count / lighter scan / warp scatter) against version 1: parity tests, then the
# compaction phase of bench.py on C2 / C3 / C4 under both.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stages.py tests/test_gpu_scale.py -x -q 2>&1 | tail -8 ) > gpurun_out/pytest_compact_v2.log
tail -3 gpurun_out/pytest_compact_v2.log
for w in c4 c3 c2; do
  for v in 2 1; do
    timeout 300 python bench.py --workload $w --no-cpu-baseline --no-wall --no-producer --steps 6 --debug-opt compact_v=$v > gpurun_out/cv_${w}_v${v}.json 2> gpurun_out/cv_${w}_v${v}.err || tail -5 gpurun_out/cv_${w}_v${v}.err
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/cv_*.json")):
    try:
        j = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    c = j["compaction"]
    print(f.split("/")[-1], "value %.4e ms/step %.3f" % (j["value"], j["ms_per_step"]), "compaction ms %.4f GB/s %.0f frac %.3f" % (c["kernel_ms"], c["achieved"], c["frac"]))
PY
