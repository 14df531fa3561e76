// dyn.cu -- batched inverse dynamics, fused trajectory + inverse dynamics, mass matrix
// and per-point forward dynamics.
//
// Replace the per-point Python loops of the reference:
//   inverse dynamics            dynamics/id_fd.py:16-48 and planning/trajectory_dynamics.py:308-380
//   gravity / Coriolis forces   dynamics/forces.py:26-133 (ddtheta = 0 / g = 0 calls)
//   mass matrix                 dynamics/mass_matrix.py:16-99
//   forward dynamics            dynamics/id_fd.py:50-83
// One thread owns one point; all link state lives in registers (mpk_device.cuh); robot
// constants are constant-bank operands.  fp64 FMA-pipe bound.
#include "mpk_common.cuh"

namespace mpk {

constexpr int kDynThreads = 128;
constexpr int kRneaMinBlocks = 5;  // 96-register cap: 20 warps / SM hide the fp64 latency

struct TipArgs {
    double g[3];
    double ftip[6];
    int has_ftip;
    const double *ftip_rows;  // (P, 6) or nullptr
};

struct RneaArgs {
    int64_t P;
    const void *th, *dth, *ddth;
    int in_dtype, vec_in;
    TipArgs tip;
    Limits lim;
    void *out;
    int out_dtype, vec_out;
};

template <int N>
__device__ __forceinline__ void store_tau(void *out, int out_dtype, bool vec, int64_t p,
                                          const double (&tau)[N], const Limits &lim) {
    if (out_dtype == MPK_F64) {
        store_row_f64<N>(static_cast<double *>(out), vec, p, tau);
    } else {
        float t32[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            t32[j] = (float)tau[j];
            if (lim.on) t32[j] = clip_f32(t32[j], lim.lo[j], lim.hi[j]);
        }
        store_row_f32<N>(static_cast<float *>(out), vec, p, t32);
    }
}

// Joint values of row p read from global memory when the recursion reaches the link.  Each
// thread walks its own contiguous row, rows of neighbouring threads are adjacent, so every
// fetched sector is fully consumed (through L1) although the individual loads are strided.
template <int N>
struct RowIn {
    const void *th, *dth, *ddth;
    int dtype;
    int64_t row;          // p * N
    double nx[3];         // joint i's values, loaded while link i - 1 was being processed
    __device__ __forceinline__ double at(const void *base, int i) const {
        if (base == nullptr) return 0.0;
        return dtype == MPK_F64 ? __ldg(static_cast<const double *>(base) + row + i)
                                : (double)__ldg(static_cast<const float *>(base) + row + i);
    }
    __device__ __forceinline__ void prefetch(int i) {
        nx[0] = at(th, i);
        nx[1] = at(dth, i);
        nx[2] = at(ddth, i);
    }
    __device__ __forceinline__ void joint(int i, double &a, double &b, double &c) {
        a = nx[0];
        b = nx[1];
        c = nx[2];
        if (i + 1 < N) prefetch(i + 1);  // overlaps the load latency with link i's arithmetic
    }
};

// Torque j of row p straight to global memory (float64, or float32 after the clip).  A thread
// fills its own contiguous row, so sectors are completed in L2 before they reach HBM.
template <int N>
struct TauOut {
    void *out;
    int dtype;
    int64_t row;
    const Limits &lim;
    __device__ __forceinline__ void put(int j, double tau) const {
        if (dtype == MPK_F64) {
            static_cast<double *>(out)[row + j] = tau;
        } else {
            float x = (float)tau;
            if (lim.on) x = clip_f32(x, lim.lo[j], lim.hi[j]);
            static_cast<float *>(out)[row + j] = x;
        }
    }
};

template <int N, bool GEN, int THREADS = kDynThreads, int MINB = kRneaMinBlocks, bool ROLLED = false>
__global__ void __launch_bounds__(THREADS, MINB)
    rnea_kernel(const __grid_constant__ RobotPack<double, N> rb, const RneaArgs a) {
    extern __shared__ __align__(16) double wsm[];
    const int64_t p = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    if (p >= a.P) return;
    double ft[6];
    const double *ftp = nullptr;
    if (a.tip.ftip_rows) {
#pragma unroll
        for (int k = 0; k < 6; ++k) ft[k] = __ldg(a.tip.ftip_rows + p * 6 + k);
        ftp = ft;
    } else if (a.tip.has_ftip) {
#pragma unroll
        for (int k = 0; k < 6; ++k) ft[k] = a.tip.ftip[k];
        ftp = ft;
    }
    RowIn<N> in{a.th, a.dth, a.ddth, a.in_dtype, p * N, {0.0, 0.0, 0.0}};
    in.prefetch(0);
    SmemStore<double, N, THREADS> st{wsm + threadIdx.x};
    if (ROLLED) {
        TauOut<N> out{a.out, a.out_dtype, p * N, a.lim};
        rnea_rolled<double, N, GEN>(rb, in, a.tip.g, ftp, st, out);
    } else {
        double tau[N];
        rnea<double, N, GEN>(rb, in, a.tip.g, ftp, tau, st);
        store_tau<N>(a.out, a.out_dtype, a.vec_out, p, tau, a.lim);
    }
}

// ---- fused trajectory + inverse dynamics ----------------------------------------
struct TrajRneaArgs {
    int64_t B, N, P;
    const double *start, *end;
    int inputs_f32;
    double Tf;
    int method;
    Limits jlim, tlim;
    TipArgs tip;
    float *tau, *pos, *vel, *acc;
    const double *ts_table;
    FastDiv div;
};

// Joint values produced from the time scaling when the recursion reaches the link: the
// float32-rounded, clipped trajectory row entries the two-call sequence would have stored.
template <int N>
struct TrajIn {
    const TrajRneaArgs &a;
    TimeScale ts;
    int64_t row;  // b * N
    __device__ __forceinline__ void joint(int i, double &th, double &qd, double &qdd) {
        double st, dth;
        endpoint(a.start, a.end, a.inputs_f32, row + i, st, dth);
        float p, v, ac;
        traj_point(ts, st, dth, a.jlim.lo[i], a.jlim.hi[i], a.jlim.on, p, v, ac);
        th = (double)p;
        qd = (double)v;
        qdd = (double)ac;
    }
};

// Torque j staged in shared memory as the float32, clipped row entry.
template <int N>
struct StageOut {
    float *row;  // staging + threadIdx.x * N
    const Limits &lim;
    __device__ __forceinline__ void put(int j, double tau) const {
        float x = (float)tau;
        if (lim.on) x = clip_f32(x, lim.lo[j], lim.hi[j]);
        row[j] = x;
    }
};

template <int N, bool GEN, int THREADS = kDynThreads, int MINB = kRneaMinBlocks, int MODE = 0>
__global__ void __launch_bounds__(THREADS, MINB)
    traj_rnea_kernel(const __grid_constant__ RobotPack<double, N> rb, const TrajRneaArgs a) {
    // dynamic shared memory: [per-thread link state of the recursion | the block's output rows,
    // staged for coalesced stores]
    extern __shared__ __align__(16) double wsm[];
    float *sm = reinterpret_cast<float *>(wsm + SmemStore<double, N, THREADS>::kSlots * 8 * THREADS);
    const int64_t p0 = (int64_t)blockIdx.x * THREADS;
    const bool live = p0 + threadIdx.x < a.P;
    int64_t b, t;
    point_coords(a.div, a.N, live ? p0 + threadIdx.x : 0, b, t);
    const int64_t rem = a.P - p0;
    const int cnt = (int)(rem < THREADS ? rem : THREADS) * N;
    const int64_t off = p0 * N;
    // (tail threads of the last block recompute point 0: they take part in every barrier and
    // their staged rows are never stored)
    TrajIn<N> in{a, time_scaling_at(a.ts_table, t, a.N, a.Tf, a.method), b * N};
    {
        double ft[6];
        const double *ftp = nullptr;
        if (a.tip.has_ftip) {
#pragma unroll
            for (int k = 0; k < 6; ++k) ft[k] = a.tip.ftip[k];
            ftp = ft;
        }
        SmemStore<double, N, THREADS> st{wsm + threadIdx.x};
        StageOut<N> out{sm + threadIdx.x * N, a.tlim};
        if (MODE == 1) {
            rnea_rolled<double, N, GEN>(rb, in, a.tip.g, ftp, st, out);
        } else {
            double tau[N];
            rnea<double, N, GEN, TrajIn<N>, SmemStore<double, N, THREADS>, MODE == 2>(rb, in, a.tip.g, ftp, tau, st);
#pragma unroll
            for (int j = 0; j < N; ++j) out.put(j, tau[j]);
        }
    }
    __syncthreads();
    tile_store(a.tau + off, sm, cnt);
}

// ---- mass matrix -------------------------------------------------------------------
struct MassArgs {
    int64_t P;
    const void *th;
    int th_dtype, vec_in, vec_out;
    double *out;
};

template <int N, bool GEN>
__global__ void __launch_bounds__(kDynThreads)
    mass_matrix_kernel(const __grid_constant__ RobotPack<double, N> rb, const MassArgs a) {
    // N^2 doubles per configuration: staged per warp and flushed coalesced (one thread writing
    // its own 8 N^2-byte row with scalar stores throttles the LSU: ncu lg_throttle 6.7)
    extern __shared__ __align__(16) double msm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *buf = msm + warp * WarpStage<N * N>::kDoubles;
    const int64_t pw = (int64_t)blockIdx.x * kDynThreads + warp * 32;
    if (pw >= a.P) return;
    const int64_t rem = a.P - pw;
    const int rows = (int)(rem < 32 ? rem : 32);
    if (lane < rows) {
        double th[N];
        load_row<N>(a.th, a.th_dtype, a.vec_in, pw + lane, th);
        JointCS<double, N> q;
        joint_cs(rb, th, q);
        double Mm[N][N];
        mass_matrix<double, N, GEN>(rb, th, q, Mm);
        double *row = buf + lane * WarpStage<N * N>::S;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) row[i * N + j] = Mm[i][j];
    }
    WarpStage<N * N>::flush(buf, a.out + pw * (N * N), rows);
}

// ---- per-point forward dynamics --------------------------------------------------------
struct FdArgs {
    int64_t P;
    const double *th, *dth, *tau;
    int vec;
    TipArgs tip;
    double *out;
};

template <int N, bool GEN>
__global__ void __launch_bounds__(kDynThreads)
    forward_dynamics_kernel(const __grid_constant__ RobotPack<double, N> rb, const FdArgs a) {
    const int64_t p = (int64_t)blockIdx.x * kDynThreads + threadIdx.x;
    if (p >= a.P) return;
    double th[N], dth[N], tau[N], dd[N];
    load_row<N>(a.th, MPK_F64, a.vec, p, th);
    load_row<N>(a.dth, MPK_F64, a.vec, p, dth);
    load_row<N>(a.tau, MPK_F64, a.vec, p, tau);
    double ft[6];
    const double *ftp = nullptr;
    if (a.tip.ftip_rows) {
#pragma unroll
        for (int k = 0; k < 6; ++k) ft[k] = __ldg(a.tip.ftip_rows + p * 6 + k);
        ftp = ft;
    } else if (a.tip.has_ftip) {
#pragma unroll
        for (int k = 0; k < 6; ++k) ft[k] = a.tip.ftip[k];
        ftp = ft;
    }
    forward_dynamics<double, N, GEN>(rb, th, dth, tau, a.tip.g, ftp, dd);
    store_row_f64<N>(a.out, a.vec, p, dd);
}

static TipArgs make_tip(const double *g, const double *Ftip, const double *Ftip_rows) {
    TipArgs t;
    for (int k = 0; k < 3; ++k) t.g[k] = g ? g[k] : 0.0;
    t.has_ftip = 0;
    for (int k = 0; k < 6; ++k) {
        t.ftip[k] = Ftip ? Ftip[k] : 0.0;
        if (t.ftip[k] != 0.0) t.has_ftip = 1;
    }
    t.ftip_rows = Ftip_rows;
    return t;
}

// Bytes of shared memory the RNEA kernels need for the per-link wrenches of one block.
template <int N>
constexpr size_t wrench_smem(int threads) {
    const size_t link_state = (size_t)(N > 1 ? N - 1 : 0) * 8 * threads * sizeof(double);
    const size_t rows = (size_t)threads * N * sizeof(float);  // output staging (fused kernel)
    return link_state + rows;
}

// Launch with dynamic shared memory, asking for the largest shared-memory carveout so that
// __launch_bounds__' blocks-per-SM target is not cut short by the L1 / shared split.
template <typename... KArgs, typename... Args>
static void launch_smem(void (*kern)(KArgs...), unsigned grid, int threads, size_t smem,
                        cudaStream_t s, Args &&...args) {
    if (smem > 32 * 1024)  // (static shared memory counts against the 48 KB default limit too)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    kern<<<grid, threads, smem, s>>>(args...);
}

static int grid_for(int64_t P, unsigned &grid) {
    const int64_t blocks = (P + kDynThreads - 1) / kDynThreads;
    if (blocks > 0x7fffffffLL) return fail(MPK_EINVAL, "point count exceeds the grid limit");
    grid = (unsigned)blocks;
    return MPK_OK;
}

}  // namespace mpk

using namespace mpk;

#define MPK_REQUIRE_DYN(rb)                                                      \
    if (!(rb)) return fail(MPK_EINVAL, "robot is NULL");                         \
    if (!(rb)->has_dynamics)                                                     \
        return fail(MPK_EINVAL, "robot was created without Glist / Mlist_per_link")

extern "C" int mpk_inverse_dynamics(const mpk_robot *rb, int64_t P, const void *theta,
                                    const void *dtheta, const void *ddtheta, int in_dtype,
                                    const double *g, const double *Ftip, const double *Ftip_rows,
                                    const float *tau_limits, void *tau, int out_dtype,
                                    void *stream) {
    MPK_REQUIRE_DYN(rb);
    if (P < 0) return fail(MPK_EINVAL, "negative size");
    if (P == 0) return MPK_OK;
    if (!theta || !tau || !g) return fail(MPK_EINVAL, "theta, g and tau are required");
    if ((in_dtype != MPK_F64 && in_dtype != MPK_F32) || (out_dtype != MPK_F64 && out_dtype != MPK_F32))
        return fail(MPK_EINVAL, "bad dtype");
    RneaArgs a;
    a.P = P;
    a.th = theta;
    a.dth = dtheta;
    a.ddth = ddtheta;
    a.in_dtype = in_dtype;
    a.vec_in = aligned16(theta) && (!dtheta || aligned16(dtheta)) && (!ddtheta || aligned16(ddtheta));
    a.tip = make_tip(g, Ftip, Ftip_rows);
    a.lim = make_limits(out_dtype == MPK_F32 ? tau_limits : nullptr, rb->n);
    a.out = tau;
    a.out_dtype = out_dtype;
    a.vec_out = aligned16(tau);
    unsigned grid;
    if (int rc = grid_for(P, grid)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (rb->rigid) {
        MPK_DISPATCH_DOF(rb->n, launch_smem(rnea_kernel<N_, false>, grid, kDynThreads, wrench_smem<N_>(kDynThreads), s, narrow<N_>(rb), a));
    } else {
        MPK_DISPATCH_DOF(rb->n, launch_smem(rnea_kernel<N_, true>, grid, kDynThreads, wrench_smem<N_>(kDynThreads), s, narrow<N_>(rb), a));
    }
    return check_launch("inverse_dynamics");
}

extern "C" int mpk_trajectory_inverse_dynamics(const mpk_robot *rb, int64_t B, int64_t N,
                                               const double *start, const double *end,
                                               int inputs_f32, double Tf, int method,
                                               const float *joint_limits, const double *g,
                                               const double *Ftip, const float *tau_limits,
                                               float *tau, float *pos, float *vel, float *acc,
                                               double *ts_scratch, void *stream) {
    MPK_REQUIRE_DYN(rb);
    if (B < 0 || N < 0) return fail(MPK_EINVAL, "negative size");
    if (B == 0 || N == 0) return MPK_OK;
    if (!start || !end || !tau || !g) return fail(MPK_EINVAL, "start, end, g and tau are required");
    for (float *o : {tau, pos, vel, acc})
        if (o && !aligned16(o)) return fail(MPK_EINVAL, "outputs must be 16-byte aligned");
    TrajRneaArgs a;
    a.B = B;
    a.N = N;
    a.P = B * N;
    a.div = make_fastdiv(N, a.P);
    a.start = start;
    a.end = end;
    a.inputs_f32 = inputs_f32;
    a.Tf = Tf;
    a.method = method;
    a.jlim = make_limits(joint_limits, rb->n);
    a.tlim = make_limits(tau_limits, rb->n);
    a.tip = make_tip(g, Ftip, nullptr);
    a.tau = tau;
    a.pos = pos;
    a.vel = vel;
    a.acc = acc;
    unsigned grid;
    if (int rc = grid_for(a.P, grid)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    a.ts_table = prepare_time_scaling(ts_scratch, B, N, Tf, method, s);
    if (pos || vel || acc) {
        // optional materialisation of the trajectory rows: the store-bound trajectory kernel
        // does it at the write-bandwidth ceiling; the fused kernel then only writes torques
        if (int rc = launch_joint_trajectory(rb->n, B, N, start, end, inputs_f32, Tf, method, joint_limits,
                                             pos, vel, acc, a.ts_table, s))
            return rc;
    }
    if (rb->rigid) {
        MPK_DISPATCH_DOF(rb->n, launch_smem(traj_rnea_kernel<N_, false>, grid, kDynThreads, wrench_smem<N_>(kDynThreads), s, narrow<N_>(rb), a));
    } else {
        MPK_DISPATCH_DOF(rb->n, launch_smem(traj_rnea_kernel<N_, true>, grid, kDynThreads, wrench_smem<N_>(kDynThreads), s, narrow<N_>(rb), a));
    }
    return check_launch("trajectory_inverse_dynamics");
}

extern "C" int mpk_mass_matrix(const mpk_robot *rb, int64_t P, const void *theta, int theta_dtype,
                               double *Mout, void *stream) {
    MPK_REQUIRE_DYN(rb);
    if (P < 0) return fail(MPK_EINVAL, "negative size");
    if (P == 0) return MPK_OK;
    if (!theta || !Mout) return fail(MPK_EINVAL, "theta and Mout are required");
    if (theta_dtype != MPK_F64 && theta_dtype != MPK_F32) return fail(MPK_EINVAL, "bad dtype");
    MassArgs a;
    a.P = P;
    a.th = theta;
    a.th_dtype = theta_dtype;
    a.vec_in = aligned16(theta);
    a.vec_out = aligned16(Mout);
    a.out = Mout;
    unsigned grid;
    if (int rc = grid_for(P, grid)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (rb->rigid) {
        MPK_DISPATCH_DOF(rb->n, launch_smem(mass_matrix_kernel<N_, false>, grid, kDynThreads,
                                            sizeof(double) * WarpStage<N_ * N_>::kDoubles * (kDynThreads / 32), s,
                                            narrow<N_>(rb), a));
    } else {
        MPK_DISPATCH_DOF(rb->n, launch_smem(mass_matrix_kernel<N_, true>, grid, kDynThreads,
                                            sizeof(double) * WarpStage<N_ * N_>::kDoubles * (kDynThreads / 32), s,
                                            narrow<N_>(rb), a));
    }
    return check_launch("mass_matrix");
}

extern "C" int mpk_forward_dynamics(const mpk_robot *rb, int64_t P, const double *theta,
                                    const double *dtheta, const double *tau, const double *g,
                                    const double *Ftip, const double *Ftip_rows, double *ddtheta,
                                    void *stream) {
    MPK_REQUIRE_DYN(rb);
    if (P < 0) return fail(MPK_EINVAL, "negative size");
    if (P == 0) return MPK_OK;
    if (!theta || !dtheta || !tau || !g || !ddtheta)
        return fail(MPK_EINVAL, "theta, dtheta, tau, g and ddtheta are required");
    FdArgs a;
    a.P = P;
    a.th = theta;
    a.dth = dtheta;
    a.tau = tau;
    a.vec = aligned16(theta) && aligned16(dtheta) && aligned16(tau) && aligned16(ddtheta);
    a.tip = make_tip(g, Ftip, Ftip_rows);
    a.out = ddtheta;
    unsigned grid;
    if (int rc = grid_for(P, grid)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (rb->rigid) {
        MPK_DISPATCH_DOF(rb->n, (forward_dynamics_kernel<N_, false><<<grid, kDynThreads, 0, s>>>(narrow<N_>(rb), a)));
    } else {
        MPK_DISPATCH_DOF(rb->n, (forward_dynamics_kernel<N_, true><<<grid, kDynThreads, 0, s>>>(narrow<N_>(rb), a)));
    }
    return check_launch("forward_dynamics");
}
