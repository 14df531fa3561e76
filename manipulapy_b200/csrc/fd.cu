// fd.cu -- forward-dynamics rollouts, parallel across trajectories, sequential in time.
//
// Replaces forward_dynamics_trajectory's CPU loop
// (planning/trajectory_dynamics.py:580-708): per row i >= 1, intRes semi-implicit Euler
// sub-steps of  ddth = M(th)^-1 (taumat[i] - c - g - Js^T Ftipmat[i]);  dth += ddth*dts;
// th += dth*dts;  th = clip(th, float32 limits);  rows are stored as float32 and the
// acceleration row is the last sub-step's.  Row 0 is the initial state with zero
// acceleration and taumat[0] is never used.  The Euler updates use explicit
// round-to-nearest multiplies and adds (no FMA contraction) like the reference's NumPy.
// One thread owns one trajectory; state, mass matrix and LDL^T factor live in registers.
#include "mpk_common.cuh"

namespace mpk {

struct RolloutArgs {
    int64_t B, N;
    const double *th0, *dth0;
    const void *taumat;
    int tau_dtype, vec_tau;
    double g[3];
    const double *ftipmat;
    double dts;
    int intRes;
    Limits lim;
    float *pos, *vel, *acc;
};

template <int N>
__device__ __forceinline__ void store_state(float *o, int64_t row, const double (&x)[N]) {
    float *r = o + row * N;
#pragma unroll
    for (int j = 0; j < N; ++j) r[j] = (float)x[j];
}

template <int N, bool GEN>
__global__ void __launch_bounds__(128)
    fd_rollout_kernel(const __grid_constant__ RobotPack<double, N> rb, const RolloutArgs a) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    double th[N], dth[N], last[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        th[j] = a.th0[b * N + j];
        dth[j] = a.dth0[b * N + j];
        last[j] = 0.0;
    }
    const int64_t base = b * a.N;
    store_state<N>(a.pos, base, th);
    store_state<N>(a.vel, base, dth);
    store_state<N>(a.acc, base, last);
    for (int64_t i = 1; i < a.N; ++i) {
        double tau[N];
        load_row<N>(a.taumat, a.tau_dtype, a.vec_tau, base + i, tau);
        double ft[6];
        const double *ftp = nullptr;
        if (a.ftipmat) {
#pragma unroll
            for (int k = 0; k < 6; ++k) ft[k] = __ldg(a.ftipmat + (base + i) * 6 + k);
            ftp = ft;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) last[j] = 0.0;
        for (int r = 0; r < a.intRes; ++r) {
            double dd[N];
            forward_dynamics<double, N, GEN>(rb, th, dth, tau, a.g, ftp, dd);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                dth[j] = rn_add(dth[j], rn_mul(dd[j], a.dts));
                double x = rn_add(th[j], rn_mul(dth[j], a.dts));
                if (a.lim.on) {
                    const double lo = (double)a.lim.lo[j], hi = (double)a.lim.hi[j];
                    x = x < lo ? lo : (x > hi ? hi : x);
                }
                th[j] = x;
                last[j] = dd[j];
            }
        }
        store_state<N>(a.pos, base + i, th);
        store_state<N>(a.vel, base + i, dth);
        store_state<N>(a.acc, base + i, last);
    }
}

}  // namespace mpk

using namespace mpk;

extern "C" int mpk_forward_dynamics_trajectory(const mpk_robot *rb, int64_t B, int64_t N,
                                               const double *theta0, const double *dtheta0,
                                               const void *taumat, int tau_dtype, const double *g,
                                               const double *Ftipmat, double dt, int intRes,
                                               const float *limits, float *pos, float *vel,
                                               float *acc, void *stream) {
    if (!rb) return fail(MPK_EINVAL, "robot is NULL");
    if (!rb->has_dynamics) return fail(MPK_EINVAL, "robot was created without Glist / Mlist_per_link");
    if (B < 0 || N < 0 || intRes < 1) return fail(MPK_EINVAL, "bad sizes");
    if (B == 0 || N == 0) return MPK_OK;
    if (!theta0 || !dtheta0 || !taumat || !g || !pos || !vel || !acc)
        return fail(MPK_EINVAL, "theta0, dtheta0, taumat, g, pos, vel, acc are required");
    if (tau_dtype != MPK_F64 && tau_dtype != MPK_F32) return fail(MPK_EINVAL, "bad dtype");
    RolloutArgs a;
    a.B = B;
    a.N = N;
    a.th0 = theta0;
    a.dth0 = dtheta0;
    a.taumat = taumat;
    a.tau_dtype = tau_dtype;
    a.vec_tau = aligned16(taumat);
    for (int k = 0; k < 3; ++k) a.g[k] = g[k];
    a.ftipmat = Ftipmat;
    a.dts = dt / (double)intRes;
    a.intRes = intRes;
    a.lim = make_limits(limits, rb->n);
    a.pos = pos;
    a.vel = vel;
    a.acc = acc;
    // few trajectories per GPU: spread them over as many SMs as possible
    int threads = 128;
    while (threads > 32 && (B + threads - 1) / threads < 2 * 148) threads >>= 1;
    const int64_t blocks = (B + threads - 1) / threads;
    if (blocks > 0x7fffffffLL) return fail(MPK_EINVAL, "B exceeds the grid limit");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (rb->rigid) {
        MPK_DISPATCH_DOF(rb->n, (fd_rollout_kernel<N_, false><<<(unsigned)blocks, threads, 0, s>>>(narrow<N_>(rb), a)));
    } else {
        MPK_DISPATCH_DOF(rb->n, (fd_rollout_kernel<N_, true><<<(unsigned)blocks, threads, 0, s>>>(narrow<N_>(rb), a)));
    }
    return check_launch("forward_dynamics_trajectory");
}
