#!/bin/bash
# Round 2: the drop-in executable -- CLI parity tests (1 and 2 GPUs), wall time of C4 as .bin / .csv.gz at -gpus 1..N
# with workers as processes (default) and as threads (-mp 0), and the bench for the compaction number.
mkdir -p gpurun_out
N=${1:-2}
( timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_abi.py -x -q 2>&1 | tail -8 ) > gpurun_out/pytest_cli.log
tail -4 gpurun_out/pytest_cli.log
python - <<PY
import json, os, subprocess, sys, time, tempfile, shutil
sys.path.insert(0, ".")
from frog_b200 import synth, build
N = $N
out = {}
for fmt in ("bin", "csv.gz"):
    tmp = tempfile.mkdtemp(prefix="fm_wall_", dir="/dev/shm")
    kps = [synth.make("iid", 20000, i) for i in range(200)]
    lst = synth.write_group(tmp, "iid", 200, 20000, fmt=fmt, threads=32, keypoints=kps)
    ref = None
    for gpus, mp in [(1, 1)] + [(g, m) for g in (2, 4, 8) if g <= N for m in (1, 0)]:
        for rep in range(2):
            t0 = time.perf_counter()
            r = subprocess.run([build.BIN, lst, "-o", tmp + "/p.bin", "-d", "1", "-d2", "0.8", "-gpus", str(gpus), "-mp", str(mp),
                                "-stats", tmp + "/s.json"], capture_output=True, text=True)
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                print("FAILED", fmt, gpus, mp, r.stderr[-500:]); break
            data = open(tmp + "/p.bin", "rb").read()
            if ref is None: ref = data
            st = json.load(open(tmp + "/s.json"))
            key = f"{fmt}_g{gpus}_mp{mp}_run{rep}"
            out[key] = dict(wall=round(wall, 3), same_bytes=data == ref, **{k: st[k] for k in ("processes", "gpu_ms_max", "pairing_s", "ctx_create_s", "upload_s", "match_call_s", "gather_s", "matches")})
            print(key, out[key], flush=True)
    shutil.rmtree(tmp)
json.dump(out, open("gpurun_out/r2_wall_c4.json", "w"), indent=1)
PY
timeout 600 python bench.py --no-cpu-baseline --no-wall --steps 5 > gpurun_out/r2c_c4_n1.json 2> gpurun_out/r2c_c4_n1.err || tail -20 gpurun_out/r2c_c4_n1.err
timeout 300 python bench.py --workload c2 --no-cpu-baseline --no-wall --steps 20 > gpurun_out/r2c_c2_n1.json 2> gpurun_out/r2c_c2_n1.err || tail -20 gpurun_out/r2c_c2_n1.err
timeout 300 python bench.py --workload c3 --no-cpu-baseline --no-wall --steps 5 > gpurun_out/r2c_c3_n1.json 2> gpurun_out/r2c_c3_n1.err || tail -20 gpurun_out/r2c_c3_n1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c_*.json")):
    j = json.load(open(f)); r = j["roofline"]
    print(f.split("/")[-1], "value %.3e e2e %.3e ms/step %.3f score_ms %.4f frac %.3f" % (j["value"], j["e2e"]["value"], j["ms_per_step"], r["kernel_ms"], r["frac"]),
          {k: round(v, 4) for k, v in j["phases_ms"].items()}, "compaction", j["compaction"])
PY
