// dyn_flavour.cu -- inverse dynamics, fused trajectory + inverse dynamics and mass-matrix
// kernels of ONE flavour (compiled three times, -DMPK_FLAVOUR=0|1|2; see dyn_kernels.cuh), general
// link geometry (GEO = 0), 1 .. 8 joints.  Flavour 0 routes the robots whose geometry signature has
// its own kernels to the geometry units (dyn_geo.cu).
#define MPK_FLAVOUR_KERNELS
#include "dyn_kernels.cuh"

#ifndef MPK_FLAVOUR
#error "compile with -DMPK_FLAVOUR=0|1|2"
#endif

namespace mpk {

template <int F>
void launch_rnea(const mpk_robot *rb, const RneaArgs &a, unsigned grid, cudaStream_t s) {
    if constexpr (F == 0) {
#define X(n_, g_) \
    if (rb->n == n_ && rb->geo == g_) return launch_rnea_n<0, n_, g_>(rb, a, grid, s);
        MPK_GEO_LIST(X)
#undef X
    }
    MPK_DISPATCH_N(launch_rnea_n, rb, a, grid, s);
}

template <int F>
void launch_traj_rnea(const mpk_robot *rb, const TrajRneaArgs &a, unsigned grid, cudaStream_t s) {
    if constexpr (F == 0) {
#define X(n_, g_) \
    if (rb->n == n_ && rb->geo == g_) return launch_traj_rnea_n<0, n_, g_>(rb, a, grid, s);
        MPK_GEO_LIST(X)
#undef X
    }
    MPK_DISPATCH_N(launch_traj_rnea_n, rb, a, grid, s);
}

template <int F>
void launch_mass(const mpk_robot *rb, const MassArgs &a, unsigned grid, cudaStream_t s) {
    if constexpr (F == 0) {
#define X(n_, g_) \
    if (rb->n == n_ && rb->geo == g_) return launch_mass_n<0, n_, g_>(rb, a, grid, s);
        MPK_GEO_LIST(X)
#undef X
    }
    MPK_DISPATCH_N(launch_mass_n, rb, a, grid, s);
}

template void launch_rnea<MPK_FLAVOUR>(const mpk_robot *, const RneaArgs &, unsigned, cudaStream_t);
template void launch_traj_rnea<MPK_FLAVOUR>(const mpk_robot *, const TrajRneaArgs &, unsigned, cudaStream_t);
template void launch_mass<MPK_FLAVOUR>(const mpk_robot *, const MassArgs &, unsigned, cudaStream_t);

}  // namespace mpk
