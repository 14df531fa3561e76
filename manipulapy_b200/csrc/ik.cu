// ik.cu -- batched damped-least-squares inverse kinematics.
//
// Replaces the per-target Python loop of SerialManipulator.iterative_inverse_kinematics
// (kinematics/ik.py:39-311), including its optional modes -- adaptive_tuning (Levenberg-Marquardt
// adaptation of the damping and the step cap, :215-229) and backtracking (five-scale line search,
// five more forward kinematics per iteration, :253-276): per
// iteration one forward kinematics + space Jacobian (kinematics/fk.py:61-70,
// jacobian.py:62-73), the geometric pose error (ik.py:88-140), the damped least-squares step
//   dtheta = V diag(s / (s^2 + lambda^2 + 1e-12)) U^T e = J^T (J J^T + (lambda^2 + 1e-12) 1)^-1 e
// (ik.py:142-162; the SVD filter and the 6 x 6 normal equations are the same map), the step cap
// and the joint-limit projection (:164-176, :253-262), best-solution tracking and the
// stagnation restart (:196-213).  One thread owns one target; the Jacobian of the current
// iterate lives in the thread's shared-memory row (odd stride, conflict free), the 6 x 6
// system is solved by LDL^T in registers.  Lanes of a warp leave the loop as they converge.
//
// Stagnation restart noise (0.1 * N(0, 1) per joint): the reference draws it from NumPy's global
// generator.  The caller may pass a table of standard normals per target (restart r uses row r) and
// read back the number of restarts taken -- the Python mirror fills it from NumPy's generator for
// single-target calls, which reproduces the reference draw for draw; without a table (batches) the
// noise comes from a counter-based generator keyed by (seed, target, iteration).
#include <cstdlib>

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
    void *queue_in;   // IkQueue this launch pops its targets from (PHASE 1)
    void *queue_out;  // IkQueue for targets still running at k_stop, or nullptr: run to the end of the budget
    int k_stop;       // iteration at which this launch hands its unfinished targets over
    const double *noise;  // (P, noise_rows, n) standard normals for the stagnation restarts, or nullptr
    int noise_rows;
    int *restarts;        // (P) restarts taken, or nullptr
};

// Queue of unfinished targets between the two phases (device workspace supplied by the caller):
// [count (8 bytes) | entries of (2 N + 8) 8-byte words: target index, best_err, (stall, k), the four
// adaptive-tuning scalars, the restart count, th, best].
template <int N>
struct IkQueue {
    static constexpr int kWords = 2 * N + 8;
    unsigned long long *count;
    double *entries;
    __device__ __forceinline__ explicit IkQueue(void *ws)
        : count(static_cast<unsigned long long *>(ws)), entries(static_cast<double *>(ws) + 2) {}
    __device__ __forceinline__ void push(int64_t target, const IkState<double, N> &st) {
        double *e = entries + atomicAdd(count, 1ULL) * kWords;
        e[0] = __longlong_as_double(target);
        e[1] = st.best_err;
        e[2] = __longlong_as_double(((long long)st.stall << 32) | (unsigned)st.k);
        e[3] = st.damping;
        e[4] = st.step_cap;
        e[5] = st.prev_err;
        e[6] = st.nu;
        e[7] = __longlong_as_double((long long)st.restarts);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            e[8 + j] = st.th[j];
            e[8 + N + j] = st.best[j];
        }
    }
    __device__ __forceinline__ int64_t pop(unsigned long long slot, IkState<double, N> &st) const {
        const double *e = entries + slot * kWords;
        st.best_err = e[1];
        const long long w = __double_as_longlong(e[2]);
        st.stall = (int)(w >> 32);
        st.k = (int)(w & 0xffffffffLL);
        st.damping = e[3];
        st.step_cap = e[4];
        st.prev_err = e[5];
        st.nu = e[6];
        st.restarts = (int)__double_as_longlong(e[7]);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            st.th[j] = e[8 + j];
            st.best[j] = e[8 + N + j];
        }
        return __double_as_longlong(e[0]);
    }
};

// PHASE 0: every target from its initial guess, up to `k_stop` iterations; with an output queue,
//          targets that are not finished by then are queued instead of keeping their warp alive.
// PHASE 1: the targets of the input queue, packed densely, up to the next `k_stop` (again queueing
//          the unfinished ones) or, in the last launch, for the rest of the budget.
// A warp lives as long as its slowest lane: on 200,000 random iiwa14 targets the mean iteration
// count is 36 but 3 % of the targets use all 400, i.e. almost every warp of a one-phase kernel
// has a lane that does.  Re-packing the survivors at geometrically growing iteration counts
// (16, 32, 64, ... ; two queues used alternately) keeps the warps full: 9.3 ms in one phase, 3.2 ms
// with one re-pack at 64, 2.5 ms with the ladder -- the floor set by the targets that run all 400
// iterations one after the other (with adaptive tuning + line search: 13.0 -> 8.7 ms); the iterates
// of every target are unchanged.
template <int N, int PHASE>
__global__ void __launch_bounds__(kIkThreads)
    ik_dls_kernel(const __grid_constant__ RobotPack<double, N> rb, const IkArgs a) {
    constexpr int S = 6 * N + 1;  // odd row stride
    extern __shared__ __align__(16) double jsm[];
    const int64_t t = (int64_t)blockIdx.x * kIkThreads + threadIdx.x;
    IkState<double, N> st;
    int64_t p;
    if (PHASE == 0) {
        if (t >= a.P) return;
        p = t;
        double th0[N];
#pragma unroll
        for (int j = 0; j < N; ++j) th0[j] = a.th0[p * N + j];
        ik_state_init(st, th0, a.prm);
    } else {
        IkQueue<N> q(a.queue_in);
        if ((unsigned long long)t >= *q.count) return;
        p = q.pop((unsigned long long)t, st);
    }
    const int k_stop = (a.queue_out && a.k_stop < a.prm.max_iter) ? a.k_stop : a.prm.max_iter;
    bool ok;
    int iters;
    const bool done = ik_dls_window<double, N>(rb, a.Td + p * 16, st, a.prm, a.seed, (unsigned long long)p,
                                               jsm + threadIdx.x * S, k_stop, ok, iters,
                                               a.noise ? a.noise + p * a.noise_rows * N : nullptr, a.noise_rows);
    if (!done) {
        IkQueue<N>(a.queue_out).push(p, st);
        return;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) a.theta[p * N + j] = st.th[j];
    a.iters[p] = iters;
    a.success[p] = ok ? 1 : 0;
    if (a.restarts) a.restarts[p] = st.restarts;
}

}  // namespace mpk

using namespace mpk;

static size_t ik_queue_bytes(int n, int64_t P) {
    return 16 + (size_t)(P > 0 ? P : 0) * (size_t)(2 * n + 8) * sizeof(double);
}

extern "C" size_t mpk_inverse_kinematics_workspace_bytes(int n, int64_t P) {
    return 2 * ik_queue_bytes(n, P);  // two queues, used alternately by the re-packing ladder
}

extern "C" int mpk_inverse_kinematics_dls(const mpk_robot *rb, int64_t P, const double *T_desired,
                                          const double *theta0, double eomg, double ev, int max_iterations,
                                          double damping, double step_cap, double weight_orientation,
                                          double weight_position, const double *joint_limits,
                                          uint64_t seed, double *theta, int32_t *iterations,
                                          uint8_t *success, void *workspace, size_t workspace_bytes,
                                          void *stream) {
    return mpk_inverse_kinematics_dls_modes(rb, P, T_desired, theta0, eomg, ev, max_iterations, damping, step_cap,
                                            weight_orientation, weight_position, joint_limits, 0, seed, nullptr, 0,
                                            theta, iterations, success, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int mpk_inverse_kinematics_dls_modes(const mpk_robot *rb, int64_t P, const double *T_desired,
                                                const double *theta0, double eomg, double ev,
                                                int max_iterations, double damping, double step_cap,
                                                double weight_orientation, double weight_position,
                                                const double *joint_limits, int flags, uint64_t seed,
                                                const double *restart_noise, int noise_rows,
                                                double *theta, int32_t *iterations, uint8_t *success,
                                                int32_t *restarts, void *workspace, size_t workspace_bytes,
                                                void *stream) {
    if (!rb) return fail(MPK_EINVAL, "robot is NULL");
    if (flags & ~(MPK_IK_ADAPTIVE_TUNING | MPK_IK_BACKTRACKING)) return fail(MPK_EINVAL, "unknown flags");
    if (P < 0 || max_iterations < 0) return fail(MPK_EINVAL, "negative size");
    if (P == 0) return MPK_OK;
    if (!T_desired || !theta0 || !theta || !iterations || !success)
        return fail(MPK_EINVAL, "T_desired, theta0, theta, iterations, success are required");
    IkArgs a;
    a.P = P;
    a.Td = T_desired;
    a.th0 = theta0;
    a.prm = make_ik_params(rb->n, eomg, ev, max_iterations, damping, step_cap, weight_orientation,
                           weight_position, joint_limits, flags);
    a.seed = seed;
    a.noise = noise_rows > 0 ? restart_noise : nullptr;
    a.noise_rows = a.noise ? noise_rows : 0;
    a.restarts = restarts;
    a.theta = theta;
    a.iters = iterations;
    a.success = success;
    const int64_t blocks = (P + kIkThreads - 1) / kIkThreads;
    if (blocks > 0x7fffffffLL) return fail(MPK_EINVAL, "P exceeds the grid limit");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // Re-packing ladder when the caller supplies a large enough workspace and there is a tail to cut.
    // First rung: 16 iterations (measured best for the plain solver, mean 36 iterations on random 7-DOF
    // targets, and with the line search, mean 16); then doubling.
    static const int first_rung = [] {  // MPK_IK_SPLIT: tuning knob, read once
        const char *e = std::getenv("MPK_IK_SPLIT");
        const int v = e ? std::atoi(e) : 0;
        return v > 0 ? v : 16;
    }();
    int rung = first_rung;
    const bool ladder = workspace && workspace_bytes >= mpk_inverse_kinematics_workspace_bytes(rb->n, P) &&
                        P >= 1024 && max_iterations > 2 * rung;
    char *queue[2] = {static_cast<char *>(workspace),
                      static_cast<char *>(workspace) + (ladder ? ik_queue_bytes(rb->n, P) : 0)};
    MPK_DISPATCH_DOF(rb->n, {
        const size_t smem = sizeof(double) * (6 * N_ + 1) * kIkThreads;
        auto k0 = ik_dls_kernel<N_, 0>;
        auto k1 = ik_dls_kernel<N_, 1>;
        if (smem > 32 * 1024) {
            cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        a.queue_in = nullptr;
        a.queue_out = nullptr;
        a.k_stop = max_iterations;
        if (!ladder) {
            k0<<<(unsigned)blocks, kIkThreads, smem, s>>>(narrow<N_>(rb), a);
        } else {
            int out = 0;
            a.queue_out = queue[out];
            a.k_stop = rung;
            cudaMemsetAsync(queue[out], 0, 16, s);
            k0<<<(unsigned)blocks, kIkThreads, smem, s>>>(narrow<N_>(rb), a);
            // (each launch is sized for the worst case; threads beyond the queued count exit at once)
            for (;;) {
                rung *= 2;
                a.queue_in = queue[out];
                out ^= 1;
                const bool last = rung >= max_iterations;
                a.queue_out = last ? nullptr : queue[out];
                a.k_stop = last ? max_iterations : rung;
                if (!last) cudaMemsetAsync(queue[out], 0, 16, s);
                k1<<<(unsigned)blocks, kIkThreads, smem, s>>>(narrow<N_>(rb), a);
                if (last) break;
            }
        }
    });
    return check_launch("inverse_kinematics_dls");
}
