#!/bin/bash
# Round evidence: smoke, parity tests, bench (ours + reference arm), ncu launch list, ncu --set full of the scoring,
# rescoring and compaction kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/gpu_info.csv 2>&1
( nproc; lscpu | grep -E "Model name|^CPU\(s\)" ) > gpurun_out/nproc.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) > gpurun_out/smoke.log
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 ) > gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 160 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^score_kernel|^rescore_kernel|^compact_scatter' -s 6 -c 3 -f -o gpurun_out/prof_final \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/smoke.log; tail -2 gpurun_out/pytest_gpu.log; cut -c1-2500 gpurun_out/bench_c2.json; cut -c1-600 gpurun_out/bench_ref.json
