// dyn_flavour.cu -- inverse dynamics, fused trajectory + inverse dynamics and mass-matrix
// kernels of ONE flavour (compiled three times, -DMPK_FLAVOUR=0|1|2; see dyn_kernels.cuh).
#define MPK_FLAVOUR_KERNELS
#include "dyn_kernels.cuh"

#ifndef MPK_FLAVOUR
#error "compile with -DMPK_FLAVOUR=0|1|2"
#endif

namespace mpk {

#define MPK_DISPATCH_DOF_V(n, ...)                               \
    switch (n) {                                                 \
        case 1: { constexpr int N_ = 1; __VA_ARGS__; } break;    \
        case 2: { constexpr int N_ = 2; __VA_ARGS__; } break;    \
        case 3: { constexpr int N_ = 3; __VA_ARGS__; } break;    \
        case 4: { constexpr int N_ = 4; __VA_ARGS__; } break;    \
        case 5: { constexpr int N_ = 5; __VA_ARGS__; } break;    \
        case 6: { constexpr int N_ = 6; __VA_ARGS__; } break;    \
        case 7: { constexpr int N_ = 7; __VA_ARGS__; } break;    \
        case 8: { constexpr int N_ = 8; __VA_ARGS__; } break;    \
        default: break;                                          \
    }

template <int F>
void launch_rnea(const mpk_robot *rb, const RneaArgs &a, unsigned grid, cudaStream_t s) {
    constexpr bool GEN = flavour_gen(F), REV = flavour_rev(F);
    if (a.compute_f32) {
        MPK_DISPATCH_DOF_V(rb->n, launch_smem_l1(rnea_kernel<float, N_, GEN, REV>, grid, kDynThreads,
                                                 wrench_smem<float, N_, GEN, REV>(), 7, s, narrow<N_, float>(rb), a));
    } else if (!GEN && !a.dth && !a.ddth && !a.tip.has_ftip) {
        // gravity forces: theta rows only, at-rest recursion (HBM-bound)
        MPK_DISPATCH_DOF_V(rb->n, launch_smem_l1(rnea_kernel<double, N_, GEN, REV, true>, grid, kDynThreads,
                                                 wrench_smem<double, N_, GEN, REV>(), 5, s, narrow<N_>(rb), a));
    } else {
        MPK_DISPATCH_DOF_V(rb->n, launch_smem_l1(rnea_kernel<double, N_, GEN, REV>, grid, kDynThreads,
                                                 wrench_smem<double, N_, GEN, REV>(), 5, s, narrow<N_>(rb), a));
    }
}

template <int F>
void launch_traj_rnea(const mpk_robot *rb, const TrajRneaArgs &a, unsigned grid, cudaStream_t s) {
    constexpr bool GEN = flavour_gen(F), REV = flavour_rev(F);
    if (a.compute_f32 && a.tip.has_ftip) {
        MPK_DISPATCH_DOF_V(rb->n, launch_smem(traj_rnea_kernel<float, N_, GEN, REV, true>, grid, kDynThreads,
                                              wrench_smem<float, N_, GEN, REV>(), s, narrow<N_, float>(rb), a));
    } else if (a.compute_f32) {
        MPK_DISPATCH_DOF_V(rb->n, launch_smem(traj_rnea_kernel<float, N_, GEN, REV, false>, grid, kDynThreads,
                                              wrench_smem<float, N_, GEN, REV>(), s, narrow<N_, float>(rb), a));
    } else if (a.pos || a.vel || a.acc) {
        // (float64, no tip wrench: the launcher only asks for this variant then)
        MPK_DISPATCH_DOF_V(rb->n, launch_smem(traj_rnea_kernel<double, N_, GEN, REV, false, true>, grid, kDynThreads,
                                              wrench_smem<double, N_, GEN, REV>() + 3 * sizeof(float) * kDynThreads * N_,
                                              s, narrow<N_>(rb), a));
    } else if (a.tip.has_ftip) {
        MPK_DISPATCH_DOF_V(rb->n, launch_smem(traj_rnea_kernel<double, N_, GEN, REV, true>, grid, kDynThreads,
                                              wrench_smem<double, N_, GEN, REV>(), s, narrow<N_>(rb), a));
    } else {
        MPK_DISPATCH_DOF_V(rb->n, launch_smem(traj_rnea_kernel<double, N_, GEN, REV, false>, grid, kDynThreads,
                                              wrench_smem<double, N_, GEN, REV>(), s, narrow<N_>(rb), a));
    }
}

template <int F>
void launch_mass(const mpk_robot *rb, const MassArgs &a, unsigned grid, cudaStream_t s) {
    constexpr bool GEN = flavour_gen(F), REV = flavour_rev(F);
    MPK_DISPATCH_DOF_V(rb->n, launch_smem(mass_matrix_kernel<N_, GEN, REV>, grid, kDynThreads,
                                          sizeof(double) * WarpStage<N_ * N_>::kDoubles * (kDynThreads / 32),
                                          s, narrow<N_>(rb), a));
}

template void launch_rnea<MPK_FLAVOUR>(const mpk_robot *, const RneaArgs &, unsigned, cudaStream_t);
template void launch_traj_rnea<MPK_FLAVOUR>(const mpk_robot *, const TrajRneaArgs &, unsigned, cudaStream_t);
template void launch_mass<MPK_FLAVOUR>(const mpk_robot *, const MassArgs &, unsigned, cudaStream_t);

}  // namespace mpk
