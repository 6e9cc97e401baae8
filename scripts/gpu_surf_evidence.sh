#!/bin/bash
# SURF3D producer on the GPU box: parity tests, bench line with the reference timed beside it, ncu of each kernel.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_surf.py -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_surf.log
tail -4 gpurun_out/pytest_surf.log
timeout 400 python scripts/gpu_surf_bench.py --tile-sweep --out gpurun_out/r2_surf_bench.json > gpurun_out/surf_bench.log 2>&1
tail -c 3000 gpurun_out/r2_surf_bench.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"response_layer|describe_kernel|integral_xy|integral_z|extrema" -c 24 \
  -o gpurun_out/r2_surf_ncu -f python scripts/gpu_surf_bench.py --steps 1 --warmup 0 --no-ref > gpurun_out/surf_ncu.log 2>&1
ncu -i gpurun_out/r2_surf_ncu.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_bytes.sum,lts__t_bytes.sum,l1tex__t_sector_hit_rate.pct,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size > gpurun_out/r2_surf_ncu_metrics.csv 2>/dev/null
wc -l gpurun_out/r2_surf_ncu_metrics.csv
