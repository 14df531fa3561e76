#!/bin/bash
# Rollout time by batch size for the single-warp and the warp-pair kernels (GPU box).
OUT=gpurun_out/fd_sizes.txt
: > $OUT
for defs in "-DMPK_FD_PAIR=0" "-DMPK_FD_PAIR=1 -DMPK_FD_PAIR_MINBLOCKS=4"; do
  echo "== $defs" >> $OUT
  MPK_FD_DEFINES="$defs" python -m manipulapy_b200._build > /dev/null 2>> $OUT
  for B in 2048 4736 9472 14208 18944 28416 37888; do
    python scripts/fd_probe.py $B 1000 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['B'], round(d['ms_min'],3))" >> $OUT
  done
done
python -m manipulapy_b200._build > /dev/null 2>&1
cat $OUT
