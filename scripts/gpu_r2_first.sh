#!/bin/bash
# Round 2, first GPU visit: parity tests, default bench (C4, wall, cpu baseline, parity sample), C2 with the
# scoring-kernel experiment variants, dense workload, launch list of one C2 step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -2
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err || tail -20 gpurun_out/r2_bench_c4_n1.err
for v in 0 1 2 3; do
  timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline --no-wall --debug-opt variant=$v > gpurun_out/r2_c2_var$v.json 2> gpurun_out/r2_c2_var$v.err || tail -5 gpurun_out/r2_c2_var$v.err
done
timeout 300 python bench.py --workload dense --steps 10 --warmup 3 --no-cpu-baseline --no-wall > gpurun_out/r2_dense.json 2> gpurun_out/r2_dense.err || tail -5 gpurun_out/r2_dense.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_*.json")):
    try:
        j = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = j["roofline"]
    print(f.split("/")[-1], "value %.3e e2e %.3e ms/step %.3f score_ms %.4f frac %.3f exec %.0f TF scored %.3f" % (
        j["value"], j["e2e"]["value"], j["ms_per_step"], r["kernel_ms"], r["frac"], r["executed_tflops"], r["scored_fraction"]),
        {k: round(v, 4) for k, v in j["phases_ms"].items()}, "compaction", round(j["compaction"]["achieved"]), "GB/s",
        "wall", j.get("wall"), "cpu", j.get("cpu_baseline"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c2.csv \
   python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-wall > gpurun_out/ncu_launch.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_c2.csv | tee gpurun_out/r2_launches_c2_summary.txt
