// fd_flavour.cu -- per-point forward dynamics and forward-dynamics rollout kernels of ONE
// flavour (compiled three times, -DMPK_FLAVOUR=0|1|2; see dyn_kernels.cuh), general link geometry
// (GEO = 0), 1 .. 8 joints.  Flavour 0 routes the robots whose geometry signature has its own
// kernels to the geometry units (fd_geo.cu).
#define MPK_FLAVOUR_KERNELS
#include "dyn_kernels.cuh"

#ifndef MPK_FLAVOUR
#error "compile with -DMPK_FLAVOUR=0|1|2"
#endif

namespace mpk {

template <int F>
void launch_fd_point(const mpk_robot *rb, const FdArgs &a, unsigned grid, cudaStream_t s) {
    if constexpr (F == 0) {
#define X(n_, g_) \
    if (rb->n == n_ && rb->geo == g_) return launch_fd_point_n<0, n_, g_>(rb, a, grid, s);
        MPK_GEO_LIST(X)
#undef X
    }
    MPK_DISPATCH_N(launch_fd_point_n, rb, a, grid, s);
}

template <int F>
void launch_rollout(const mpk_robot *rb, const RolloutArgs &a, cudaStream_t s) {
    if constexpr (F == 0) {
#define X(n_, g_) \
    if (rb->n == n_ && rb->geo == g_) return launch_rollout_n<0, n_, g_>(rb, a, s);
        MPK_GEO_LIST(X)
#undef X
    }
    MPK_DISPATCH_N(launch_rollout_n, rb, a, s);
}

template void launch_fd_point<MPK_FLAVOUR>(const mpk_robot *, const FdArgs &, unsigned, cudaStream_t);
template void launch_rollout<MPK_FLAVOUR>(const mpk_robot *, const RolloutArgs &, cudaStream_t);

}  // namespace mpk
