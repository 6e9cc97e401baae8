#!/bin/bash
# bench at N=1 and N=2 on a 2-GPU box (no CPU baseline)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python - <<'PY'
import json
for n in (1,2):
    try:
        j=json.load(open(f"gpurun_out/bench_n{n}.json"))
        print("N=%d value %.4e (%.3f ms)  e2e %.4e (%.3f ms)  score_ms %.4f frac %.3f" % (n, j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["roofline"]["kernel_ms"], j["roofline"]["frac"]))
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/bench_n{n}.err").read()[-1500:])
PY
