#!/bin/bash
# experiment: time the fused kernel under each launch variant (MPK_VARIANT), device time only
OUT=gpurun_out; mkdir -p $OUT
for v in ${1:-0 1 2 3 4 5 6 7 8 9}; do
  MPK_VARIANT=$v timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('variant $v', 'ms', round(d['ms_per_step'],4), 'pts/s %.3e'%d['value'], 'fp64frac', round(d['roofline']['fp64']['frac'],3))"
done | tee $OUT/variants.txt
