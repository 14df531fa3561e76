// ik.cu -- batched damped-least-squares inverse kinematics.
//
// Replaces the per-target Python loop of SerialManipulator.iterative_inverse_kinematics
// (kinematics/ik.py:39-311) in its default mode (adaptive_tuning = backtracking = False): per
// iteration one forward kinematics + space Jacobian (kinematics/fk.py:61-70,
// jacobian.py:62-73), the geometric pose error (ik.py:88-140), the damped least-squares step
//   dtheta = V diag(s / (s^2 + lambda^2 + 1e-12)) U^T e = J^T (J J^T + (lambda^2 + 1e-12) 1)^-1 e
// (ik.py:142-162; the SVD filter and the 6 x 6 normal equations are the same map), the step cap
// and the joint-limit projection (:164-176, :253-262), best-solution tracking and the
// stagnation restart (:196-213).  One thread owns one target; the Jacobian of the current
// iterate lives in the thread's shared-memory row (odd stride, conflict free), the 6 x 6
// system is solved by LDL^T in registers.  Lanes of a warp leave the loop as they converge.
//
// Deviation: the stagnation restart adds 0.1 * N(0, 1) noise; the reference draws it from
// NumPy's global generator, this kernel from a counter-based generator keyed by (seed, target,
// iteration).  Runs that never stagnate for 20 iterations -- the usual case -- do not touch it.
#include "mpk_common.cuh"

namespace mpk {

constexpr int kIkThreads = 128;

struct IkArgs {
    int64_t P;
    const double *Td, *th0;
    IkParams<double, MPK_MAX_DOF> prm;
    unsigned long long seed;
    double *theta;
    int *iters;
    unsigned char *success;
};

template <int N>
__global__ void __launch_bounds__(kIkThreads)
    ik_dls_kernel(const __grid_constant__ RobotPack<double, N> rb, const IkArgs a) {
    constexpr int S = 6 * N + 1;  // odd row stride
    extern __shared__ __align__(16) double jsm[];
    const int64_t p = (int64_t)blockIdx.x * kIkThreads + threadIdx.x;
    if (p >= a.P) return;
    double th[N];
#pragma unroll
    for (int j = 0; j < N; ++j) th[j] = a.th0[p * N + j];
    int iters;
    const bool ok = ik_dls<double, N>(rb, a.Td + p * 16, th, a.prm, a.seed, (unsigned long long)p,
                                      jsm + threadIdx.x * S, iters);
#pragma unroll
    for (int j = 0; j < N; ++j) a.theta[p * N + j] = th[j];
    a.iters[p] = iters;
    a.success[p] = ok ? 1 : 0;
}

}  // namespace mpk

using namespace mpk;

extern "C" int mpk_inverse_kinematics_dls(const mpk_robot *rb, int64_t P, const double *T_desired,
                                          const double *theta0, double eomg, double ev, int max_iterations,
                                          double damping, double step_cap, double weight_orientation,
                                          double weight_position, const double *joint_limits,
                                          uint64_t seed, double *theta, int32_t *iterations,
                                          uint8_t *success, void *stream) {
    if (!rb) return fail(MPK_EINVAL, "robot is NULL");
    if (P < 0 || max_iterations < 0) return fail(MPK_EINVAL, "negative size");
    if (P == 0) return MPK_OK;
    if (!T_desired || !theta0 || !theta || !iterations || !success)
        return fail(MPK_EINVAL, "T_desired, theta0, theta, iterations, success are required");
    IkArgs a;
    a.P = P;
    a.Td = T_desired;
    a.th0 = theta0;
    a.prm = make_ik_params(rb->n, eomg, ev, max_iterations, damping, step_cap, weight_orientation,
                           weight_position, joint_limits);
    a.seed = seed;
    a.theta = theta;
    a.iters = iterations;
    a.success = success;
    const int64_t blocks = (P + kIkThreads - 1) / kIkThreads;
    if (blocks > 0x7fffffffLL) return fail(MPK_EINVAL, "P exceeds the grid limit");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MPK_DISPATCH_DOF(rb->n, {
        const size_t smem = sizeof(double) * (6 * N_ + 1) * kIkThreads;
        auto kern = ik_dls_kernel<N_>;
        if (smem > 32 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<(unsigned)blocks, kIkThreads, smem, s>>>(narrow<N_>(rb), a);
    });
    return check_launch("inverse_kinematics_dls");
}
