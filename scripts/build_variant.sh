#!/bin/bash
# Build a tuning variant of the native libraries next to the product build:
#   scripts/build_variant.sh <name> "<defines for the dyn units>" ["<defines for the fd units>" ["<defines for traj.cu>"]]
# -> manipulapy_b200/_lib_<name>/{libmpk.so,_mpk_ops.so}; run anything with MPK_LIB_DIR pointing there.
# The product build in manipulapy_b200/_lib is restored afterwards.
set -e
cd "$(dirname "$0")/.."
name=$1
rm -rf manipulapy_b200/_lib_keep && cp -r manipulapy_b200/_lib manipulapy_b200/_lib_keep
MPK_DYN_DEFINES="$2" MPK_FD_DEFINES="$3" MPK_TRAJ_DEFINES="$4" python -m manipulapy_b200._build > /dev/null
rm -rf manipulapy_b200/_lib_$name && mkdir manipulapy_b200/_lib_$name
cp manipulapy_b200/_lib/libmpk.so manipulapy_b200/_lib/_mpk_ops.so manipulapy_b200/_lib_$name/
for f in manipulapy_b200/_lib/obj/dyn_geo_6_d52ad0.ptxas.log; do cp $f manipulapy_b200/_lib_$name/; done
rm -rf manipulapy_b200/_lib && mv manipulapy_b200/_lib_keep manipulapy_b200/_lib
echo "built manipulapy_b200/_lib_$name"
