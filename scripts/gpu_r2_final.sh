#!/bin/bash
# Round-2 closing evidence on one B200: whole GPU test suite, smoke(), both bench arms, the producer's bench and ncu.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/r2_pytest_gpu_final.log
tail -3 gpurun_out/r2_pytest_gpu_final.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r2_smoke_final.log
tail -1 gpurun_out/r2_smoke_final.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_final.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/r2_bench_c4_final2.json 2> gpurun_out/bench_c4.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2_bench_c4_final2.json"))
print("c4 value %.4e e2e %.4e ms/step %.1f frac %.3f" % (j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["frac"]))
print("wall", json.dumps(j.get("wall"))[:600])
print("producer", json.dumps(j.get("producer"))[:900])
print("cpu_baseline", json.dumps(j.get("cpu_baseline"))[:500])
r = json.load(open("gpurun_out/r2_bench_reference_final.json"))
print("reference arm value %.4e" % r["value"])
PY
[ -n "$SKIP_SURF" ] || timeout 400 python scripts/gpu_surf_bench.py --tile-sweep --out gpurun_out/r2_surf_bench_final.json > gpurun_out/surf_bench.log 2>&1
[ -n "$SKIP_SURF" ] || timeout 400 ncu --set full --clock-control none --import-source on -k regex:"response_layer|describe_kernel|integral_xy|integral_z|extrema|interpolate" -c 24 \
  -o gpurun_out/r2_surf_ncu_final -f python scripts/gpu_surf_bench.py --steps 1 --warmup 0 --no-ref > gpurun_out/surf_ncu.log 2>&1
ncu -i gpurun_out/r2_surf_ncu_final.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size > gpurun_out/r2_surf_ncu_final_metrics.csv 2>/dev/null
wc -l gpurun_out/r2_surf_ncu_final_metrics.csv
