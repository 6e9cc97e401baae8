#!/bin/bash
# N = 1 and N = 2 strong-scaling bench lines (no wall / cpu legs) + pinned H2D bandwidth of the box
mkdir -p gpurun_out
N=${1:-2}
python - <<'PY'
import torch, time
a = torch.empty(1 << 28, dtype=torch.uint8).pin_memory(); b = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); b.copy_(a, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print("pinned H2D 256 MiB: %.1f GB/s" % (a.numel() / dt / 1e9))
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); a.copy_(b, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print("pinned D2H 256 MiB: %.1f GB/s" % (a.numel() / dt / 1e9))
PY
timeout 600 python bench.py --no-cpu-baseline --no-wall --steps 6 > gpurun_out/r2h_c4_n1.json 2> gpurun_out/r2h_c4_n1.err || tail -20 gpurun_out/r2h_c4_n1.err
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530+n)) \
     bench.py --gpus $n --steps 6 --no-wall > gpurun_out/r2h_c4_n$n.json 2> gpurun_out/r2h_c4_n$n.err || tail -30 gpurun_out/r2h_c4_n$n.err
done
python - <<'PY'
import json, glob
base = None
for f in sorted(glob.glob("gpurun_out/r2h_*.json"), key=lambda s: int(s.split("_n")[-1].split(".")[0])):
    try:
        j = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    if j["n_gpus"] == 1: base = j
    eff = (j["value"] / base["value"] / j["n_gpus"], j["e2e"]["value"] / base["e2e"]["value"] / j["n_gpus"]) if base else (0, 0)
    print(f.split("/")[-1], "N", j["n_gpus"], "value %.3e e2e %.3e ms/step %.2f e2e ms %.2f eff value %.3f e2e %.3f" % (
        j["value"], j["e2e"]["value"], j["ms_per_step"], j["e2e"]["ms_per_step"], eff[0], eff[1]),
        "h2d", j["e2e"]["h2d_bytes_per_step"], "host", j["rank0_host_ms_per_step"], j["e2e"]["rank0_host_ms_per_step"], j["phases_ms"])
PY
