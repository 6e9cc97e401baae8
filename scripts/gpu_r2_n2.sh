#!/bin/bash
# Round 2: two GPUs -- 2-GPU CLI parity tests, strong-scaling bench at N = 2 (C4), and N = 1 for the ratio.
mkdir -p gpurun_out
N=${1:-2}
( timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpus or compaction or distances" 2>&1 | tail -8 ) > gpurun_out/pytest_two.log
tail -4 gpurun_out/pytest_two.log
timeout 900 python bench.py --no-cpu-baseline --steps 6 > gpurun_out/r2b_c4_n1.json 2> gpurun_out/r2b_c4_n1.err || tail -20 gpurun_out/r2b_c4_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 6 > gpurun_out/r2b_c4_n$N.json 2> gpurun_out/r2b_c4_n$N.err || tail -30 gpurun_out/r2b_c4_n$N.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_*.json")):
    try:
        j = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = j["roofline"]
    print(f.split("/")[-1], "N", j["n_gpus"], "value %.3e e2e %.3e ms/step %.3f e2e ms %.3f score_ms %.4f frac %.3f" % (
        j["value"], j["e2e"]["value"], j["ms_per_step"], j["e2e"]["ms_per_step"], r["kernel_ms"], r["frac"]),
        {k: round(v, 4) for k, v in j["phases_ms"].items()}, "compaction", round(j["compaction"]["achieved"]), "GB/s",
        "host", j["rank0_host_ms_per_step"], j["e2e"]["rank0_host_ms_per_step"], "wall", j.get("wall"))
PY
