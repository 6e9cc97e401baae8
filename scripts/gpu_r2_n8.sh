#!/bin/bash
# Round 2, N-GPU box: strong-scaling bench of ONE C4 group at N GPUs (and N = 1 on the same box), then the wall time of
# the drop-in executable on C4 / C3 (.bin) one-shot and served, and the CUDA start-up probe.
mkdir -p gpurun_out
N=${1:-8}
nproc; nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
timeout 600 python bench.py --no-cpu-baseline --no-wall --steps 6 > gpurun_out/r2s_c4_n1.json 2> gpurun_out/r2s_c4_n1.err || tail -20 gpurun_out/r2s_c4_n1.err
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) \
     bench.py --gpus $n --steps 6 --no-wall > gpurun_out/r2s_c4_n$n.json 2> gpurun_out/r2s_c4_n$n.err || tail -30 gpurun_out/r2s_c4_n$n.err
done
python - <<'PY'
import json, glob
base = None
for f in sorted(glob.glob("gpurun_out/r2s_*.json"), key=lambda s: int(s.split("_n")[-1].split(".")[0])):
    try:
        j = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    if j["n_gpus"] == 1: base = j
    eff = (j["value"] / base["value"] / j["n_gpus"], j["e2e"]["value"] / base["e2e"]["value"] / j["n_gpus"]) if base else (0, 0)
    print(f.split("/")[-1], "N", j["n_gpus"], "value %.3e e2e %.3e ms/step %.2f e2e ms %.2f eff value %.3f e2e %.3f" % (
        j["value"], j["e2e"]["value"], j["ms_per_step"], j["e2e"]["ms_per_step"], eff[0], eff[1]),
        "host", j["rank0_host_ms_per_step"], j["e2e"]["rank0_host_ms_per_step"])
PY
python - <<PY
import json, os, subprocess, sys, time, tempfile, shutil
sys.path.insert(0, ".")
from frog_b200 import synth, build
N = $N
out = {}
def run(lst, tmp, extra, env):
    t0 = time.perf_counter()
    r = subprocess.run([build.BIN, lst, "-o", tmp + "/p.bin", "-stats", tmp + "/s.json"] + extra, capture_output=True, text=True, env=env)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        return {"failed": r.stderr[-300:]}
    st = json.load(open(tmp + "/s.json"))
    return dict(wall=round(wall, 3), **{k: st[k] for k in ("gpus", "processes", "gpu_ms_max", "pairing_s", "ctx_create_s", "upload_s", "match_call_s", "matches")})
for name, n_img, n_pts, flags in (("c4", 200, 20000, ["-d", "1", "-d2", "0.8"]), ("c3", 50, 50000, ["-d", "1"])):
    tmp = tempfile.mkdtemp(prefix="fm_wall_", dir="/dev/shm")
    kps = [synth.make("iid", n_pts, i) for i in range(n_img)]
    lst = synth.write_group(tmp, "iid", n_img, n_pts, fmt="bin", threads=32, keypoints=kps)
    del kps
    env = dict(os.environ, FROGMATCH_SOCKET=tmp + "/fm.sock", FROGMATCH_SERVE_IDLE="300")
    ref = None
    for gpus, mp in [(1, 0), (N, 0), (N, 1), (0, 0)]:
        tag = f"{name}_g{gpus if gpus else 'auto'}_mp{mp}"
        extra = flags + (["-gpus", str(gpus)] if gpus else []) + ["-mp", str(mp)]
        runs = [run(lst, tmp, extra, env) for _ in range(2)]
        data = open(tmp + "/p.bin", "rb").read()
        ref = ref or data
        out[tag] = dict(runs=runs, same_bytes=data == ref)
        print(tag, out[tag], flush=True)
    # served: the first call starts the server (contexts on all N GPUs)
    start = run(lst, tmp, flags + ["-gpus", str(N), "-serve", "1"], env)
    for gpus in (1, N, 0):
        tag = f"{name}_g{gpus if gpus else 'auto'}_served"
        extra = flags + (["-gpus", str(gpus)] if gpus else []) + ["-serve", "1"]
        runs = [run(lst, tmp, extra, env) for _ in range(2)]
        out[tag] = dict(runs=runs, same_bytes=open(tmp + "/p.bin", "rb").read() == ref, server_start_call=start)
        print(tag, out[tag], flush=True)
    subprocess.run([build.BIN, "-serve-stop"], env=env)
    shutil.rmtree(tmp)
json.dump(out, open(f"gpurun_out/r2_wall_box{N}.json", "w"), indent=1)
PY
bash scripts/gpu_ctx_probe.sh $N > /dev/null 2>&1; cat gpurun_out/r2_ctx_probe_n$N.txt | tail -30
