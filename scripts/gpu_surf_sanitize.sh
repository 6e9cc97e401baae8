#!/bin/bash
# compute-sanitizer over the SURF3D producer: memcheck, racecheck, synccheck, initcheck on smoke_surf() (golden "small"
# volume: every kernel of libfrogsurf.so runs) and memcheck + racecheck on the descriptor / voxel-type tests.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  timeout 600 $CS --tool $tool --print-limit 20 --log-file gpurun_out/r2_sanitizer_${tool}_surf_smoke.txt python -c "import __graft_entry__ as g; g.smoke_surf()" > gpurun_out/san_${tool}_surf_smoke.out 2>&1
  echo "$tool surf smoke: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_${tool}_surf_smoke.txt | tail -1)"
done
for tool in memcheck racecheck; do
  timeout 900 $CS --tool $tool --print-limit 20 --log-file gpurun_out/r2_sanitizer_${tool}_surf_tests.txt python -m pytest tests/test_gpu_surf.py -x -q -k "descriptors_on_reference_points or voxel_types or stages_match_golden and f32" > gpurun_out/san_${tool}_surf_tests.out 2>&1
  echo "$tool surf tests: rc=$? $(tail -1 gpurun_out/san_${tool}_surf_tests.out) $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_${tool}_surf_tests.txt | tail -1)"
done
