// fd_flavour.cu -- per-point forward dynamics and forward-dynamics rollout kernels of ONE
// flavour (compiled three times, -DMPK_FLAVOUR=0|1|2; see dyn_kernels.cuh).
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
void launch_fd_point(const mpk_robot *rb, const FdArgs &a, unsigned grid, cudaStream_t s) {
    constexpr bool GEN = flavour_gen(F), REV = flavour_rev(F);
    MPK_DISPATCH_DOF_V(rb->n, (forward_dynamics_kernel<N_, GEN, REV><<<grid, kDynThreads, 0, s>>>(narrow<N_>(rb), a)));
}

static int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 1;
    }
    return cached[dev];
}

// One warp per block by default: the kernel needs no block-level cooperation, and single-warp
// blocks spread a small batch over all SMs (8,192 rollouts: 4.96 ms against 5.7 ms with 128-thread
// blocks; 65,536 rollouts: no difference).
template <int F>
void launch_rollout(const mpk_robot *rb, const RolloutArgs &a, cudaStream_t s) {
    constexpr bool GEN = flavour_gen(F), REV = flavour_rev(F);
    constexpr int threads = kRolloutThreads;
    const unsigned grid = (unsigned)((a.B + threads - 1) / threads);
#if MPK_FD_PAIR
    // A batch that fits the GPU in one wave of warp pairs (4 blocks x 32 rollouts per SM) runs each
    // step split across two warps: a lone warp is bound by its own instruction issue (~1950 fp64
    // instructions at one per two cycles on ONE scheduler's fp64 unit), the pair uses two schedulers.
    // Measured (iiwa14, 1000 steps): 2,048 rollouts 3.30 -> 2.76 ms, 8,192: 3.95 -> 3.25, 18,944:
    // 4.71 -> 4.00; beyond one wave the single-warp kernel wins (28,416: 5.60 against 7.13 ms).
    if (!GEN && REV && !a.ftipmat && rb->n >= 2 && a.B <= (int64_t)kRolloutPairBlocksPerSm * 32 * sm_count()) {
        // plain revolute chain, rigid links, no tip wrench
        MPK_DISPATCH_DOF_V(rb->n, launch_smem(fd_rollout_pair_kernel<N_>, (unsigned)((a.B + 31) / 32), 64,
                                              rollout_pair_smem<N_>(), s, narrow<N_>(rb), a));
        return;
    }
#endif
    if (a.ftipmat) {
        MPK_DISPATCH_DOF_V(rb->n, launch_smem(fd_rollout_kernel<N_, GEN, REV, true>, grid, threads,
                                              rollout_smem_per_warp<N_>() * (threads / 32), s, narrow<N_>(rb), a));
    } else {
        MPK_DISPATCH_DOF_V(rb->n, launch_smem(fd_rollout_kernel<N_, GEN, REV, false>, grid, threads,
                                              rollout_smem_per_warp<N_>() * (threads / 32), s, narrow<N_>(rb), a));
    }
}

template void launch_fd_point<MPK_FLAVOUR>(const mpk_robot *, const FdArgs &, unsigned, cudaStream_t);
template void launch_rollout<MPK_FLAVOUR>(const mpk_robot *, const RolloutArgs &, cudaStream_t);

}  // namespace mpk
