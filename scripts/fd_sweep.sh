#!/bin/bash
# Rebuild the forward-dynamics units with each set of defines and time the rollouts (on the GPU box).
#   gpurun -- 'bash scripts/fd_sweep.sh tag "-DMPK_FD_MINBLOCKS=12" "-DMPK_FD_MINBLOCKS=16"'
TAG=$1; shift
OUT=gpurun_out/${TAG}_sweep.txt
: > $OUT
for defs in "$@"; do
  echo "== $defs" >> $OUT
  MPK_FD_DEFINES="$defs" python -m manipulapy_b200._build > /dev/null 2>> $OUT
  grep -A2 "fd_rollout.*ILi7ELb0ELb1ELb0" manipulapy_b200/_lib/obj/fd_flavour0.ptxas.log | grep -E "spill|Used" >> $OUT
  python scripts/fd_probe.py 65536 1000 3 >> $OUT 2>/dev/null
  python scripts/fd_probe.py 8192 1000 3 >> $OUT 2>/dev/null
done
python -m manipulapy_b200._build > /dev/null 2>&1
cat $OUT
