// fd_geo.cu -- the forward-dynamics kernels (per point and rollouts) of ONE link-geometry signature of
// plain revolute chains (compiled once per entry of MPK_GEO_LIST, -DMPK_GEO_N=<joints>
// -DMPK_GEO_SIG=<signature>; see "link geometry classes" in mpk_device.cuh).
#define MPK_FLAVOUR_KERNELS
#define MPK_GEO_UNIT
#include "dyn_kernels.cuh"

#if !defined(MPK_GEO_N) || !defined(MPK_GEO_SIG)
#error "compile with -DMPK_GEO_N=<joints> -DMPK_GEO_SIG=<signature>"
#endif

namespace mpk {
template void launch_fd_point_n<0, MPK_GEO_N, MPK_GEO_SIG>(const mpk_robot *, const FdArgs &, unsigned, cudaStream_t);
template void launch_rollout_n<0, MPK_GEO_N, MPK_GEO_SIG>(const mpk_robot *, const RolloutArgs &, cudaStream_t);
}  // namespace mpk
