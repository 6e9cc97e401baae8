#!/bin/bash
# Round 2: whole GPU test-suite (with the at-scale parity tests), then the default bench and C3 / C2 lines.
mkdir -p gpurun_out
T0=$(date +%s)
( timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 ) > gpurun_out/pytest_gpu_full.log
tail -16 gpurun_out/pytest_gpu_full.log
echo "pytest took $(( $(date +%s) - T0 )) s"
timeout 900 python bench.py > gpurun_out/r2d_c4_n1.json 2> gpurun_out/r2d_c4_n1.err || tail -20 gpurun_out/r2d_c4_n1.err
timeout 300 python bench.py --workload c3 --no-cpu-baseline --no-wall --steps 5 > gpurun_out/r2d_c3_n1.json 2> gpurun_out/r2d_c3_n1.err || tail -20 gpurun_out/r2d_c3_n1.err
timeout 300 python bench.py --workload c2 --no-cpu-baseline --no-wall --steps 20 > gpurun_out/r2d_c2_n1.json 2> gpurun_out/r2d_c2_n1.err || tail -20 gpurun_out/r2d_c2_n1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2d_*.json")):
    try:
        j = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = j["roofline"]
    print(f.split("/")[-1], "value %.3e e2e %.3e ms/step %.3f score_ms %.4f frac %.3f" % (j["value"], j["e2e"]["value"], j["ms_per_step"], r["kernel_ms"], r["frac"]),
          {k: round(v, 4) for k, v in j["phases_ms"].items()}, "compaction GB/s", round(j["compaction"]["achieved"]), "rej", round(j.get("rows_rejected_early_frac", -1), 4),
          "cand/row", round(j["candidates_per_row"], 3), "\n   wall", j.get("wall"), "\n   cpu", j.get("cpu_baseline"))
PY
