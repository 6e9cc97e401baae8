#!/bin/bash
# Strong scaling of bin/match on G GPUs at the big BASELINE configs (reference run kept tiny: first 2 images)
G=${1:-8}
mkdir -p gpurun_out
for cfg in c3 c4; do
  timeout 600 python scripts/scale_check.py --config $cfg --gpus $G --sub 2 > gpurun_out/scale_${cfg}_g$G.json 2> gpurun_out/scale_${cfg}_g$G.err
  echo "scale $cfg rc=$?"; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/scale_${cfg}_g$G.json"))
    print({k: j.get(k) for k in ("match_wall_s","phases_s","matches","ref_blocks_compared","ref_blocks_mismatching","blocks_violating_properties","gpu_pairs_per_s_kernels","ok")}, j.get("stats"))
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/scale_${cfg}_g$G.err").read()[-1500:])
PY
done
