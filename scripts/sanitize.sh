#!/bin/bash
# compute-sanitizer passes over small invocations of every kernel family (GPU box):
# memcheck on a slice of the GPU parity tests, racecheck on the split rollout kernels
# (shared-memory hand-over with named barriers) and the staged-output kernels.
OUT=gpurun_out/sanitize.txt
: > $OUT
echo "== memcheck: pytest -k 'golden or zoo and ur5 or rollout_kernels_agree or modes'" >> $OUT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q \
    -k "dynamics_golden or joint_trajectory_golden or (zoo and ur3) or batched_rollouts or inverse_kinematics_modes or cartesian" 2>&1 | tail -6 >> $OUT
echo "== racecheck: rollouts (three-warp kernel 2048 x 30 [one group per block] and 6000 x 30 [two groups], pair kernel 12000 x 12, single-warp kernel 24000 x 6, pair kernel at 6 blocks per SM 40000 x 6), fused trajectory + inverse dynamics" >> $OUT
for cfg in "2048 30" "6000 30" "12000 12" "24000 6" "40000 6"; do
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/fd_probe.py $cfg 1 2>&1 | tail -2 | cut -c1-200 >> $OUT
done
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "fused_trajectory_inverse_dynamics_equals_two_calls" 2>&1 | tail -4 >> $OUT
cat $OUT
