// dyn_geo.cu -- the inverse-dynamics, fused and mass-matrix kernels of ONE link-geometry signature
// of plain revolute chains (compiled once per entry of MPK_GEO_LIST, -DMPK_GEO_N=<joints>
// -DMPK_GEO_SIG=<signature>; see "link geometry classes" in mpk_device.cuh).
#define MPK_FLAVOUR_KERNELS
#define MPK_GEO_UNIT
#include "dyn_kernels.cuh"

#if !defined(MPK_GEO_N) || !defined(MPK_GEO_SIG)
#error "compile with -DMPK_GEO_N=<joints> -DMPK_GEO_SIG=<signature>"
#endif

namespace mpk {
template void launch_rnea_n<0, MPK_GEO_N, MPK_GEO_SIG>(const mpk_robot *, const RneaArgs &, unsigned, cudaStream_t);
template void launch_traj_rnea_n<0, MPK_GEO_N, MPK_GEO_SIG>(const mpk_robot *, const TrajRneaArgs &, unsigned, cudaStream_t);
template void launch_mass_n<0, MPK_GEO_N, MPK_GEO_SIG>(const mpk_robot *, const MassArgs &, unsigned, cudaStream_t);
}  // namespace mpk
