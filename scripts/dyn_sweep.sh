#!/bin/bash
# Rebuild the inverse-dynamics units with each set of defines and time the fused kernel (GPU box).
TAG=$1; shift
OUT=gpurun_out/${TAG}_dynsweep.txt
: > $OUT
for defs in "$@"; do
  echo "== $defs" >> $OUT
  MPK_DYN_DEFINES="$defs" python -m manipulapy_b200._build > /dev/null 2>> $OUT
  grep -A2 "traj_rnea_kernelIdLi6ELb0ELb1ELb0ELb0" manipulapy_b200/_lib/obj/dyn_flavour0.ptxas.log | grep -E "spill|Used" >> $OUT
  python bench.py --steps 10 --warmup 3 --no-cpu --no-fd 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms', round(d['ms_per_step'],4), 'pts/s %.4g'%d['value'], 'pipe', round(d['roofline']['fp64']['pipe_issue_frac'],3))" >> $OUT
done
python -m manipulapy_b200._build > /dev/null 2>&1
cat $OUT
