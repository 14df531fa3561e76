#!/bin/bash
# bench.py at N = $@ GPUs of one box, one JSON line each:  gpurun --gpus 8 -- 'bash scripts/bench_scale.sh r2t 8 4 2'
tag=$1; shift
mkdir -p gpurun_out/$tag
for N in "$@"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/$tag/bench_n$N.json 2> gpurun_out/$tag/bench_n$N.err
  tail -2 gpurun_out/$tag/bench_n$N.err | cut -c1-200
done
