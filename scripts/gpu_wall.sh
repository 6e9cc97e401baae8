#!/bin/bash
# Wall time of the one-shot executable on the box's GPUs at C4: -gpus 1 / all / default (planned from file sizes)
mkdir -p gpurun_out /tmp/fmw
python - <<'PY'
import sys, os, time, subprocess, json, filecmp
sys.path.insert(0, os.getcwd())
from frog_b200 import synth
lst = synth.write_group("/tmp/fmw/c4", "iid", 200, 20000, fmt="bin")
ng = len(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.strip().splitlines())
res = {}
for g in ("1", str(ng), "default", "1"):
    out, st = f"/tmp/fmw/out_{g}.bin", f"/tmp/fmw/st_{g}.json"
    cmd = ["./bin/match", lst, "-o", out, "-d", "1", "-d2", "0.8", "-stats", st] + ([] if g == "default" else ["-gpus", g])
    t0 = time.time(); r = subprocess.run(cmd, capture_output=True, text=True); wall = time.time() - t0
    s = json.load(open(st))
    res[g] = dict(wall_s=round(wall, 3), rc=r.returncode, gpus=s["gpus"], pairing_s=s["pairing_s"], ctx_create_s=s["ctx_create_s"],
                  upload_s=s["upload_s"], match_call_s=s["match_call_s"], gpu_ms_max=s["gpu_ms_max"])
    print(g, res[g], flush=True)
res["identical"] = filecmp.cmp("/tmp/fmw/out_1.bin", f"/tmp/fmw/out_{ng}.bin", False) and filecmp.cmp("/tmp/fmw/out_1.bin", "/tmp/fmw/out_default.bin", False)
res["box_gpus"] = ng
json.dump(res, open("gpurun_out/wall_c4.json", "w"), indent=1)
print("identical:", res["identical"])
PY
