#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for cfg in "c3 3" "c4 4"; do
  set -- $cfg
  timeout 1200 python scripts/scale_check.py --config $1 --gpus 1 --sub $2 > gpurun_out/scale_$1_g1.json 2> gpurun_out/scale_$1_g1.err
  echo "scale $1 rc=$?"; python -c "
import json;j=json.load(open('gpurun_out/scale_$1_g1.json'));print({k:j[k] for k in ('ok','match_wall_s','phases_s','ref_blocks_mismatching')}, {k:j['stats'][k] for k in ('gpu_ms_max','ctx_create_s','upload_s','match_call_s')})"
done
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cut -c1-700 gpurun_out/bench_c2.json
