#!/bin/bash
# Round 2: compaction check (bench lines), then compute-sanitizer memcheck / racecheck / synccheck / initcheck over
# smoke() and the group parity test (small inputs: the tools slow kernels down 10-100x).
mkdir -p gpurun_out
for w in c2 c3 c4; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline --no-wall --steps 5 > gpurun_out/r2e_${w}_n1.json 2> gpurun_out/r2e_${w}_n1.err || tail -20 gpurun_out/r2e_${w}_n1.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2e_*.json")):
    j = json.load(open(f)); r = j["roofline"]
    print(f.split("/")[-1], "value %.3e ms/step %.3f score_ms %.4f" % (j["value"], j["ms_per_step"], r["kernel_ms"]),
          {k: round(v, 4) for k, v in j["phases_ms"].items()}, "compaction", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in j["compaction"].items()})
PY
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 $CS --tool $tool --print-limit 20 --log-file gpurun_out/r2_sanitizer_${tool}_smoke.txt python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_${tool}_smoke.out 2>&1
  echo "$tool smoke: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_${tool}_smoke.txt | tail -1)"
done
for tool in memcheck racecheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 --log-file gpurun_out/r2_sanitizer_${tool}_parity.txt python -m pytest tests/test_gpu_parity.py -x -q -k "test_group_parity and bank and 0.22 or test_sym_and_repeated or test_distances_side_output" > gpurun_out/san_${tool}_parity.out 2>&1
  echo "$tool parity: rc=$? $(tail -1 gpurun_out/san_${tool}_parity.out) $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_${tool}_parity.txt | tail -1)"
done
