#!/bin/bash
# Rollout time by batch size, per kernel choice (GPU box): MPK_FD_SPLIT = 1 single-warp kernel only, 2 = no
# three-warp kernel, 3 = the launcher's own choice (three warps / two warps / one warp / pair at 6 blocks per SM),
# 4 = pair kernel (4 blocks per SM) whatever the batch.
OUT=gpurun_out/fd_sizes.txt
: > $OUT
for K in 1 2 3 4; do
  echo "== MPK_FD_SPLIT=$K" >> $OUT
  for B in 2048 4736 8192 9472 14208 18944 28416 37888 40000 65536 131072; do
    MPK_FD_SPLIT=$K python scripts/fd_probe.py $B 1000 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['B'], round(d['ms_min'],3))" >> $OUT
  done
done
cat $OUT
