#!/bin/bash
# One GPU box visit: parity tests, bench, ncu launch list, ncu full capture of the scoring kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/gpu_info.csv 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 2 -c 2 -o gpurun_out/prof_score \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_c2.json; cat gpurun_out/bench_ref.json
