#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both modes), ncu launch list and a full
# capture of the dominant kernel.  Usage (from the build container):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag] [sections]'
# sections: any of  tests smoke bench launches ncu ncu_micro micro

set -u
TAG=${1:-r1}
SECTIONS=${2:-"tests smoke bench launches ncu"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
for s in $SECTIONS; do
case $s in
tests)
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log ;;
smoke)
  timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1
  tail -2 $OUT/${TAG}_smoke.log ;;
bench)
  timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
  timeout 300 python bench.py --steps 20 --warmup 3 --mode two_kernel --no-cpu > $OUT/${TAG}_bench_two_kernel.json 2>> $OUT/${TAG}_bench.err
  cat $OUT/${TAG}_bench_two_kernel.json
  timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
  cat $OUT/${TAG}_bench_reference.json ;;
launches)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-fd > $OUT/${TAG}_launches.log 2>&1
  tail -2 $OUT/${TAG}_launches.log ;;
ncu)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:traj_rnea_kernel -s 2 -c 2 \
      -f -o $OUT/${TAG}_traj_rnea python bench.py --steps 2 --warmup 3 --no-cpu --no-fd > $OUT/${TAG}_ncu.log 2>&1
  tail -2 $OUT/${TAG}_ncu.log ;;
ncu_micro)
  # full captures of the other kernels (one launch each) while running the micro-benchmark
  # (no --import-source here: 24 full captures with source exceed gpurun's 64 MiB return limit)
  timeout 1200 ncu --set full --clock-control none \
      -k regex:'traj_kernel|fk_jacobian_kernel|fd_rollout|mass_matrix_kernel|rnea_kernel|ik_dls_kernel|forward_dynamics_kernel' -c 24 \
      -f -o $OUT/${TAG}_micro python scripts/microbench.py --quick > $OUT/${TAG}_ncu_micro.log 2>&1
  tail -2 $OUT/${TAG}_ncu_micro.log
  # summarise on the box and drop the report: 24 full captures are ~90 MB, gpurun returns <= 64 MiB
  python scripts/ncu_summary.py $OUT/${TAG}_micro.ncu-rep --md $OUT/${TAG}_ncu_kernels_micro.md > /dev/null 2>&1
  rm -f $OUT/${TAG}_micro.ncu-rep ;;
micro)
  timeout 900 python scripts/microbench.py > $OUT/${TAG}_micro.json 2> $OUT/${TAG}_micro.err
  cat $OUT/${TAG}_micro.json; tail -3 $OUT/${TAG}_micro.err ;;
esac
done
