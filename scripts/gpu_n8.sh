#!/bin/bash
# bench at N GPUs (arg 1) under torchrun, as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"
python - <<PY
import json
try:
    t=open("gpurun_out/bench_n$N.json").read().strip().splitlines()
    print("stdout lines:", len(t))
    j=json.loads(t[-1])
    print(j["e2e"].get("rank0_host_ms_per_step")); print("N=%d value %.4e (%.3f ms)  e2e %.4e (%.3f ms)  score_ms %.4f" % (j["n_gpus"], j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["kernel_ms"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_n$N.err").read()[-2500:])
PY
