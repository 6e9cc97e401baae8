#!/bin/bash
# 2-GPU box: the CLI's two-GPU tests (host and NCCL gather), then bench at N=2 under torchrun
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_cli.py -m gpu -q -k "two_gpus" 2>&1 | tail -4 ) > gpurun_out/pytest_two.log; cat gpurun_out/pytest_two.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python - <<'PY'
import json
t=open("gpurun_out/bench_n2.json").read().strip().splitlines()
j=json.loads(t[-1]); print("stdout lines", len(t), "N=2 value %.4e (%.3f ms) e2e %.4e (%.3f ms) score %.4f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["kernel_ms"]))
PY
