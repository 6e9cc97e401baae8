#!/bin/bash
# BASELINE.json configs end to end through bin/match on G GPUs (default 1), checked against the reference binary.
G=${1:-1}
mkdir -p gpurun_out
for cfg in "c1 0" "c2b 0" "c3 5" "c4 8"; do
  set -- $cfg
  timeout 1200 python scripts/scale_check.py --config $1 --gpus $G --sub $2 > gpurun_out/scale_$1_g$G.json 2> gpurun_out/scale_$1_g$G.err
  echo "scale $1 rc=$?"; cut -c1-900 gpurun_out/scale_$1_g$G.json; tail -3 gpurun_out/scale_$1_g$G.err
done
if [ "$G" = "1" ]; then
  for w in c3 c4; do
    timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
    echo "bench $w rc=$?"; cut -c1-1500 gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
  done
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/bench_c2_g$G.json 2> gpurun_out/bench_c2_g$G.err
  echo "bench c2 x$G rc=$?"; cut -c1-1500 gpurun_out/bench_c2_g$G.json; tail -3 gpurun_out/bench_c2_g$G.err
fi
