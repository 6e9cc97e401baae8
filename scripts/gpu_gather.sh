#!/bin/bash
# bin/match on G GPUs: match lists gathered over NCCL to GPU 0 vs fetched per GPU, checked against the reference binary
G=${1:-2}
mkdir -p gpurun_out
for cfg in "c1 0" "c2b 0" "c3 5"; do
  set -- $cfg
  for mode in nccl host; do
    timeout 900 python scripts/scale_check.py --config $1 --gpus $G --sub $2 --gather $mode > gpurun_out/gather_$1_${mode}_g$G.json 2> gpurun_out/gather_$1_${mode}_g$G.err
    echo "gather $1 $mode rc=$?"; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/gather_$1_${mode}_g$G.json"))
    print({k: j.get(k) for k in ("match_wall_s","phases_s","matches","identical_blocks","checked_blocks","ok")}, j.get("stats"))
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/gather_$1_${mode}_g$G.err").read()[-1500:])
PY
  done
done
