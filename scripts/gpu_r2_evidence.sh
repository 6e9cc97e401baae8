#!/bin/bash
# Round 2 evidence set on one B200: smoke, whole GPU suite, both bench arms on the default workload (C4), C2 / C3 /
# dense lines, launch list of a C2 step, ncu --set full of the scoring / rescoring / compaction kernels on C2.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( timeout 1800 python -m pytest tests -m gpu -q --durations=6 2>&1 | tail -14 ) > gpurun_out/r2_pytest_gpu.log; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err || tail -5 gpurun_out/r2_bench_reference_arm.err
timeout 900 python bench.py > gpurun_out/r2_bench_c4_n1_final.json 2> gpurun_out/r2_bench_c4_n1_final.err || tail -20 gpurun_out/r2_bench_c4_n1_final.err
for w in c2 c3 dense; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --no-wall --steps 10 > gpurun_out/r2_bench_${w}_n1_final.json 2> gpurun_out/r2_bench_${w}_n1_final.err || tail -5 gpurun_out/r2_bench_${w}_n1_final.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*final.json")) + ["gpurun_out/r2_bench_reference_arm.json"]:
    try:
        j = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = j.get("roofline", {})
    print(f.split("/")[-1], "value %.3e e2e %.3e ms/step %.3f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]),
          "score_ms %.3f frac %.3f exec %.0f TF (%.3f)" % (r.get("kernel_ms", 0), r.get("frac", 0), r.get("executed_tflops", 0), r.get("executed_frac", 0)) if r else "",
          j.get("phases_ms"), "compaction", j.get("compaction", {}).get("achieved"), "\n    wall", j.get("wall"), "\n    cpu", j.get("cpu_baseline"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c2.csv \
   python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-wall > gpurun_out/ncu_launch.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_c2.csv | tee gpurun_out/r2_launches_c2_summary.txt | head -12
for k in score_kernel rescore_kernel compact_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k -s 3 -c 1 -f -o gpurun_out/r2_ncu_c2_$k \
     python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-wall > gpurun_out/ncu_c2_$k.log 2>&1
  tail -1 gpurun_out/ncu_c2_$k.log | cut -c1-120
done
