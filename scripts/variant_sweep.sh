#!/bin/bash
# Time the headline step with every tuning variant built by scripts/build_variant.sh:
#   scripts/variant_sweep.sh <outdir> [bench flags]
out=$1; shift
mkdir -p $out
for d in manipulapy_b200/_lib manipulapy_b200/_lib_*; do
  [ -f $d/libmpk.so ] || continue
  n=$(basename $d)
  MPK_LIB_DIR=$d python bench.py --no-configs --no-sweep --no-cpu --no-fd "$@" > $out/bench_$n.json 2> $out/bench_$n.err
  python -c "
import json,sys; d=json.loads(open('$out/bench_$n.json').read()); print('$n', round(d['ms_per_step'],4), '%.4e' % d['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
