// dyn_kernels.cuh -- kernels and argument blocks of the dynamics launchers.
//
// Replace the per-point Python loops of the reference:
//   inverse dynamics            dynamics/id_fd.py:16-48 and planning/trajectory_dynamics.py:308-380
//   gravity / Coriolis forces   dynamics/forces.py:26-133 (ddtheta = 0 / g = 0 calls)
//   mass matrix                 dynamics/mass_matrix.py:16-99
//   forward dynamics            dynamics/id_fd.py:50-83
//   forward-dynamics rollouts   planning/trajectory_dynamics.py:580-708
// One thread owns one point (or one rollout); robot constants are constant-bank operands
// (mpk_device.cuh).  fp64 FMA-pipe bound.
//
// Every kernel exists in three FLAVOURS, each compiled in its own translation unit
// (dyn_flavour.cu / fd_flavour.cu with -DMPK_FLAVOUR=k) so that the build parallelises:
//   0  rigid link inertias, all joints revolute, plain D-H links   (UR5, iiwa, ...: most URDF arms)
//   1  rigid link inertias, a prismatic joint or a Hayati link past a revolute first joint
//      (the reference's 8-DOF Panda)
//   2  general symmetric 6x6 link inertias (and the rare rigid chain whose first joint is prismatic)
#pragma once
#include <cstdlib>

#include "mpk_common.cuh"

namespace mpk {

constexpr int kDynThreads = 128;
// 4 blocks (16 warps) per SM, 127-register cap.  Measured on B200 (profiles/r1_variants.md): 4, 5
// and 6 resident blocks run the fused 6-DOF kernel within 2 % of each other -- the fp64 pipe, not
// latency hiding, is the limit -- and the looser register cap gives the shortest code.
#ifndef MPK_RNEA_MINBLOCKS
#define MPK_RNEA_MINBLOCKS 4
#endif
constexpr int kRneaMinBlocks = MPK_RNEA_MINBLOCKS;
// The fused kernel: 6 resident blocks (24 warps) per SM -- what its shared memory allows -- i.e. an
// 85-register cap.  Its prologue evaluates the sines / cosines of all joints as independent chains,
// which the compiler would otherwise spread over 120 registers (4 blocks per SM: fp64 pipe 67 % busy,
// top stalls `wait` and `barrier`; with 6 blocks and no block barrier: see profiles/r2_variants.md).
#ifndef MPK_FUSED_MINBLOCKS
#define MPK_FUSED_MINBLOCKS 6
#endif
constexpr int kFusedMinBlocks = MPK_FUSED_MINBLOCKS;
// MPK_RNEA_REGSTORE (tuning knob): the fused kernel keeps the link wrenches in registers (RegStore)
// instead of the per-thread shared-memory column.
#ifndef MPK_RNEA_REGSTORE
#define MPK_RNEA_REGSTORE 0
#endif

constexpr bool flavour_gen(int f) { return f == 2; }
constexpr bool flavour_rev(int f) { return f == 0; }
inline int flavour_of(const mpk_robot *rb) {
    return (!rb->rigid || !rb->first_revolute) ? 2 : (rb->plain ? 0 : 1);
}

struct TipArgs {
    double g0[3];  // -g in frame-0 coordinates (base_gravity)
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
    int compute_f32;  // run the recursion in float32 (north-star tolerance 1e-4 on torques)
};

struct TrajRneaArgs {
    int64_t B, N, P;
    const double *start, *end;
    int inputs_f32;
    double Tf;
    int method;
    Limits jlim, tlim;
    TipArgs tip;
    float *tau;
    float *pos, *vel, *acc;  // only read by the WRITE variant of the fused kernel
    const double *ts_table;
    FastDiv div;
    int compute_f32;
    int table_rows;  // trajectories one thread group's consecutive points can touch (traj_table_rows)
};

struct MassArgs {
    int64_t P;
    const void *th;
    int th_dtype, vec_in, vec_out;
    double *out;
};

struct FdArgs {
    int64_t P;
    const double *th, *dth, *tau;
    int vec;
    TipArgs tip;
    double *out;
};

struct RolloutArgs {
    int64_t B, N;
    const double *th0, *dth0;
    const void *taumat;
    int tau_dtype, vec_tau;
    double g0[3];
    const double *ftipmat;
    double dts;
    int intRes;
    Limits lim;
    float *pos, *vel, *acc;
};

// Flavour launchers: defined and explicitly instantiated in dyn_flavour.cu / fd_flavour.cu.
// Rollout kernel launch shape (tuning knobs, profiles/r1_variants.md E): MPK_FD_THREADS threads per
// block, MPK_FD_MINBLOCKS resident blocks per SM (the register cap: 16,384 / (32 x warps per
// scheduler) registers per thread), MPK_FD_PHASES block barriers per Euler step.
#ifndef MPK_FD_THREADS
#define MPK_FD_THREADS 32
#endif
#ifndef MPK_FD_MINBLOCKS
#define MPK_FD_MINBLOCKS (256 / MPK_FD_THREADS)
#endif
#ifndef MPK_FD_PHASES
#define MPK_FD_PHASES 0
#endif
// MPK_FD_PAIR: plain revolute chains take fd_rollout_pair_kernel (a step split across two warps)
#ifndef MPK_FD_PAIR
#define MPK_FD_PAIR 1
#endif
constexpr int kRolloutThreads = MPK_FD_THREADS;

template <int FLAVOUR> void launch_rnea(const mpk_robot *rb, const RneaArgs &a, unsigned grid, cudaStream_t s);
template <int FLAVOUR> void launch_traj_rnea(const mpk_robot *rb, const TrajRneaArgs &a, unsigned grid, cudaStream_t s);
template <int FLAVOUR> void launch_mass(const mpk_robot *rb, const MassArgs &a, unsigned grid, cudaStream_t s);
template <int FLAVOUR> void launch_fd_point(const mpk_robot *rb, const FdArgs &a, unsigned grid, cudaStream_t s);
template <int FLAVOUR> void launch_rollout(const mpk_robot *rb, const RolloutArgs &a, cudaStream_t s);

// per joint count (and geometry signature): defined under MPK_FLAVOUR_KERNELS below, instantiated by
// the flavour units (GEO = 0, N = 1..8) and the geometry units (flavour 0, one (N, GEO) each)
template <int F, int N, unsigned GEO> void launch_rnea_n(const mpk_robot *rb, const RneaArgs &a, unsigned grid, cudaStream_t s);
template <int F, int N, unsigned GEO> void launch_traj_rnea_n(const mpk_robot *rb, const TrajRneaArgs &a, unsigned grid, cudaStream_t s);
template <int F, int N, unsigned GEO> void launch_mass_n(const mpk_robot *rb, const MassArgs &a, unsigned grid, cudaStream_t s);
template <int F, int N, unsigned GEO> void launch_fd_point_n(const mpk_robot *rb, const FdArgs &a, unsigned grid, cudaStream_t s);
template <int F, int N, unsigned GEO> void launch_rollout_n(const mpk_robot *rb, const RolloutArgs &a, cudaStream_t s);

#define MPK_DISPATCH_FLAVOUR(rb, CALL)                 \
    switch (flavour_of(rb)) {                          \
        case 0: { constexpr int F_ = 0; CALL; } break; \
        case 1: { constexpr int F_ = 1; CALL; } break; \
        default: { constexpr int F_ = 2; CALL; } break; \
    }

// Bytes of shared memory the RNEA kernels need per block: the per-link wrenches of the
// recursion; the fused kernel re-uses the same bytes to stage its output rows.
template <typename T, int N, bool GEN, bool REV>
constexpr size_t wrench_smem() {
    const size_t link_state = SmemStore<T, N, kDynThreads, rnea_fast0(GEN, REV, N)>::kBytes;
    const size_t rows = (size_t)kDynThreads * N * sizeof(float);
    return link_state > rows ? link_state : rows;
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

// The same, but carving out only the shared memory `blocks` resident blocks need, so that the
// rest of the 256 KB stays L1: kernels that read their input rows with strided per-joint loads
// rely on L1 to hold a warp's rows between the loads of consecutive joints.
template <typename... KArgs, typename... Args>
static void launch_smem_l1(void (*kern)(KArgs...), unsigned grid, int threads, size_t smem, int blocks,
                           cudaStream_t s, Args &&...args) {
    if (smem > 32 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const size_t need = (smem + 1024) * (size_t)blocks;
    int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    kern<<<grid, threads, smem, s>>>(args...);
}

// Threads that share one shared-memory slice of the fused kernel (link state / output staging / endpoint
// table) and synchronise among themselves: a warp (32, only __syncwarp) or the whole block (128,
// __syncthreads).  Measured on B200 (profiles/r2_variants.md): MPK_FUSED_GROUP.
#ifndef MPK_FUSED_GROUP
#define MPK_FUSED_GROUP 128
#endif
constexpr int kFusedGroup = MPK_FUSED_GROUP;
static_assert(kFusedGroup == 32 || kFusedGroup == kDynThreads, "a group is a warp or the block");

// shared memory of the endpoint tables: (start, delta) x N joints x the trajectories one group's
// consecutive points can touch, one table per group
__host__ __device__ inline int traj_table_rows(int64_t B, int64_t N) {
    int64_t k = (kFusedGroup - 1) / (N > 0 ? N : 1) + 2;
    return (int)(k > B ? B : k);
}
inline size_t traj_table_bytes(int n, int64_t B, int64_t N) {
    return (size_t)(kDynThreads / kFusedGroup) * traj_table_rows(B, N) * n * 2 * sizeof(double);
}

#ifdef MPK_FLAVOUR_KERNELS
// ======================================================================================
// kernels (only the flavour translation units see this part)
// ======================================================================================

template <int N, typename T>
__device__ __forceinline__ void store_tau(void *out, int out_dtype, bool vec, int64_t p,
                                          const T (&tau_t)[N], const Limits &lim) {
    if (out_dtype == MPK_F64) {
        double tau[N];
#pragma unroll
        for (int j = 0; j < N; ++j) tau[j] = (double)tau_t[j];
        store_row_f64<N>(static_cast<double *>(out), vec, p, tau);
    } else {
        const T (&tau)[N] = tau_t;
        float t32[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            t32[j] = (float)tau[j];
            if (lim.on) t32[j] = clip_f32(t32[j], lim.lo[j], lim.hi[j]);
        }
        store_row_f32<N>(static_cast<float *>(out), vec, p, t32);
    }
}

__device__ __forceinline__ const double *load_tip(const TipArgs &tip, int64_t p, double (&ft)[6]) {
    if (tip.ftip_rows) {
#pragma unroll
        for (int k = 0; k < 6; ++k) ft[k] = __ldg(tip.ftip_rows + p * 6 + k);
        return ft;
    }
    if (tip.has_ftip) {
#pragma unroll
        for (int k = 0; k < 6; ++k) ft[k] = tip.ftip[k];
        return ft;
    }
    return nullptr;
}

// Joint values of row p.  float64 rows are read from global memory when the recursion reaches
// the link, one joint ahead (each thread walks its own contiguous row, rows of neighbouring
// threads are adjacent, so every fetched sector is consumed through L1 -- the launcher leaves
// L1 room for that).  float32 rows (the trajectory-level API: 3 x 4 N bytes per point) are
// fetched whole up front with 8 / 16-byte vector loads and wait in 3 N 32-bit registers: all
// loads of a row are in flight together and every sector is requested at most twice.
// REST: the rows of dtheta and ddtheta are absent (gravity forces) -- only theta is read and the
// recursion takes its at-rest form.
template <int N, typename T, bool REST = false>
struct RowIn {
    static constexpr bool kZeroAcc = REST;
    static constexpr bool kZeroVel = REST;
    const void *th, *dth, *ddth;
    int dtype;
    int64_t row;          // p * N
    T nx[3];              // float64 rows: joint i's values, loaded while link i - 1 was being processed
    float r32[3][N];      // float32 rows
    __device__ __forceinline__ T at(const void *base, int i) const {
        if (base == nullptr) return T(0);
        return (T)__ldg(static_cast<const double *>(base) + row + i);
    }
    __device__ __forceinline__ void prefetch(int i) {
        nx[0] = at(th, i);
        if (!REST) {
            nx[1] = at(dth, i);
            nx[2] = at(ddth, i);
        }
    }
    __device__ __forceinline__ void load_f32(const void *base, float (&dst)[N], bool vec) {
        if (base == nullptr) {
#pragma unroll
            for (int j = 0; j < N; ++j) dst[j] = 0.f;
            return;
        }
        const float *g = static_cast<const float *>(base) + row;
        if (vec && N % 4 == 0) {
#pragma unroll
            for (int j = 0; j < N / 4; ++j) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(g) + j);
                dst[4 * j] = v.x; dst[4 * j + 1] = v.y; dst[4 * j + 2] = v.z; dst[4 * j + 3] = v.w;
            }
        } else if (vec && N % 2 == 0) {
#pragma unroll
            for (int j = 0; j < N / 2; ++j) {
                const float2 v = __ldg(reinterpret_cast<const float2 *>(g) + j);
                dst[2 * j] = v.x; dst[2 * j + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < N; ++j) dst[j] = __ldg(g + j);
        }
    }
    __device__ __forceinline__ void begin(bool vec) {
        if (dtype == MPK_F64) {
            prefetch(0);
        } else {
            load_f32(th, r32[0], vec);
            if (!REST) {
                load_f32(dth, r32[1], vec);
                load_f32(ddth, r32[2], vec);
            }
        }
    }
    __device__ __forceinline__ void joint(int i, T &a, T &b, T &c) {
        if (dtype == MPK_F64) {
            a = nx[0];
            b = REST ? T(0) : nx[1];
            c = REST ? T(0) : nx[2];
            if (i + 1 < N) prefetch(i + 1);  // overlaps the load latency with link i's arithmetic
        } else {
            a = (T)r32[0][i];
            b = REST ? T(0) : (T)r32[1][i];
            c = REST ? T(0) : (T)r32[2][i];
        }
    }
};

// gravity / tip wrench of the argument block in the kernel's arithmetic type
template <typename T>
struct TipT {
    T g0[3];
    T ft[6];
};
template <typename T>
__device__ __forceinline__ const T *tip_to(const TipArgs &tip, const double *ftp, TipT<T> &o) {
#pragma unroll
    for (int k = 0; k < 3; ++k) o.g0[k] = (T)tip.g0[k];
    if (!ftp) return nullptr;
#pragma unroll
    for (int k = 0; k < 6; ++k) o.ft[k] = (T)ftp[k];
    return o.ft;
}

template <typename T, int N, bool GEN, bool REV, bool REST = false, unsigned GEO = 0>
__global__ void __launch_bounds__(kDynThreads, kRneaMinBlocks)
    rnea_kernel(const __grid_constant__ RobotPack<T, N> rb, const RneaArgs a) {
    extern __shared__ __align__(16) double wsm_raw[];
    T *wsm = reinterpret_cast<T *>(wsm_raw);
    const int64_t p = (int64_t)blockIdx.x * kDynThreads + threadIdx.x;
    if (p >= a.P) return;
    double ft[6];
    const double *ftp = load_tip(a.tip, p, ft);
    TipT<T> tt;
    const T *ftt = tip_to<T>(a.tip, ftp, tt);
    RowIn<N, T, REST> in{a.th, a.dth, a.ddth, a.in_dtype, p * N, {T(0), T(0), T(0)}, {}};
    in.begin(a.vec_in != 0);
    SmemStore<T, N, kDynThreads, rnea_fast0(GEN, REV, N)> st{wsm + 2 * threadIdx.x};
    T tau[N];
    rnea<T, N, GEN, REV, GEO>(rb, in, tt.g0, ftt, tau, st);
    store_tau<N, T>(a.out, a.out_dtype, a.vec_out, p, tau, a.lim);
}

// ---- fused trajectory + inverse dynamics ----------------------------------------
// The trajectory rows are never read back from HBM: each thread produces the float32-rounded,
// clipped row entries the two-call sequence would have stored and feeds them to the recursion.
//
// Prologue (everything that is not rigid-body algebra, kept as short as it can be -- the kernel is
// bound by the fp64 pipe AND the register-file bandwidth, which every instruction shares):
//   1. the block's 128 consecutive points belong to at most 127 / N + 2 trajectories: their
//      endpoints (start, end - start, in the precision the reference subtracts in) are staged ONCE
//      per block in shared memory, so a thread's per-joint work is two 8-byte shared loads instead
//      of two global loads, three conversions and a subtraction;
//   2. all joint positions first, then all joint sines / cosines behind ONE range test: N
//      independent dependency chains in one basic block, coefficients fetched once (SmemStorePre);
//   3. velocities / accelerations of a joint are produced when the recursion reaches its link.
// `stage` (WRITE variant): the thread's row in the block's shared-memory tiles of positions,
// velocities and accelerations, from where the rows go to HBM with coalesced stores.
template <int N, typename T>
struct TrajInPre {
    double sd, sdd;
    const double *tab;  // this thread's trajectory in the block's endpoint table: (start, delta) per joint
    float *stage;       // or nullptr
    __device__ __forceinline__ void joint(int i, T &th, T &qd, T &qdd) {
        const double dth = tab[2 * i + 1];
        const float v = (float)rn_mul(sd, dth), ac = (float)rn_mul(sdd, dth);
        if (stage) {
            stage[kDynThreads * N + i] = v;
            stage[2 * kDynThreads * N + i] = ac;
        }
        th = T(0);  // (the joint rotations were evaluated up front)
        qd = (T)v;
        qdd = (T)ac;
    }
};

// The same, positions included, everything produced when the recursion reaches the link (the joint
// rotation is then evaluated there too: MPK_FUSED_LAZY_SINCOS, fewer registers, one range test and
// one coefficient fetch per link).
template <int N, typename T>
struct TrajInLazy {
    const TimeScale &ts;
    const double *tab;
    const Limits &jlim;
    float *stage;
    __device__ __forceinline__ void joint(int i, T &th, T &qd, T &qdd) {
        const double st = tab[2 * i], dth = tab[2 * i + 1];
        const float pj = clip_f32((float)rn_add(rn_mul(ts.s, dth), st), jlim.lo[i], jlim.hi[i]);
        const float v = (float)rn_mul(ts.sd, dth), ac = (float)rn_mul(ts.sdd, dth);
        if (stage) {
            stage[i] = pj;
            stage[kDynThreads * N + i] = v;
            stage[2 * kDynThreads * N + i] = ac;
        }
        th = (T)pj;
        qd = (T)v;
        qdd = (T)ac;
    }
};
#ifndef MPK_FUSED_LAZY_SINCOS
#define MPK_FUSED_LAZY_SINCOS 0
#endif

// WRITE: also materialise the trajectory rows (positions, velocities, accelerations).  The
// kernel is bound by the fp64 pipe with HBM at 6 %, so the extra 12 N bytes per point ride
// along at a fraction of their stand-alone cost (0.65 ms against 0.17 + 0.57 ms for the two launches).
__device__ __forceinline__ void fused_group_sync() {
    if (kFusedGroup == 32) __syncwarp();
    else __syncthreads();
}

template <typename T, int N, bool GEN, bool REV, bool TIP, bool WRITE = false, unsigned GEO = 0>
__global__ void __launch_bounds__(kDynThreads, kFusedMinBlocks)
    traj_rnea_kernel(const __grid_constant__ RobotPack<T, N> rb, const TrajRneaArgs a) {
    // Dynamic shared memory, one slice per GROUP of kFusedGroup threads:
    //   the per-thread link state of the recursion (column stride = group size), whose bytes afterwards
    //   stage the group's output rows for coalesced stores | the group's endpoint table;
    //   WRITE: + three block-wide trajectory tiles.
    extern __shared__ __align__(16) double wsm_raw[];
    constexpr int G = kFusedGroup, kGroups = kDynThreads / G;
#if MPK_FUSED_LAZY_SINCOS
    using Store = SmemStore<T, N, G, rnea_fast0(GEN, REV, N)>;
#else
    using Store = SmemStorePre<T, N, G, rnea_fast0(GEN, REV, N)>;
#endif
    constexpr size_t kGroupState = wrench_smem<T, N, GEN, REV>() / kGroups;  // bytes, a multiple of 128
    const int grp = threadIdx.x / G, gl = threadIdx.x % G;
    char *smem = reinterpret_cast<char *>(wsm_raw);
    T *wsm = reinterpret_cast<T *>(smem + grp * kGroupState);
    float *sm = reinterpret_cast<float *>(smem + grp * kGroupState);
    float *traj_sm = reinterpret_cast<float *>(smem + wrench_smem<T, N, GEN, REV>());
    const int table_rows = a.table_rows;
    double *table = reinterpret_cast<double *>(smem + wrench_smem<T, N, GEN, REV>() +
                                               (WRITE ? 3 * sizeof(float) * kDynThreads * N : 0)) +
                    grp * table_rows * (2 * N);
    const int64_t pw = (int64_t)blockIdx.x * kDynThreads + grp * G;  // the group's first point
    if (G == 32 && !WRITE && pw >= a.P) return;                        // whole warp past the end
    const int64_t rem = a.P - pw;
    const int live_n = rem <= 0 ? 0 : (int)(rem < G ? rem : G);
    // (threads past the end recompute the group's first point; their staged rows are never stored.  In
    // the WRITE variant a warp past the end idles through the block barrier on point 0.)
    const int64_t pf = live_n > 0 ? pw : 0;
    int64_t b, t, b0, t0, b1, t1;
    point_coords(a.div, a.N, gl < live_n ? pw + gl : pf, b, t);
    point_coords(a.div, a.N, pf, b0, t0);
    point_coords(a.div, a.N, live_n > 0 ? pw + live_n - 1 : pf, b1, t1);
    {
        const int entries = (int)(b1 - b0 + 1) * N;
        for (int e = gl; e < entries; e += G) {
            double st, dth;
            endpoint(a.start, a.end, a.inputs_f32, b0 * N + e, st, dth);
            table[2 * e] = st;
            table[2 * e + 1] = dth;
        }
    }
    fused_group_sync();
    const TimeScale ts = time_scaling_at(a.ts_table, t, a.N, a.Tf, a.method);
    const double *tab = table + (b - b0) * (2 * N);
    float *stage = WRITE ? traj_sm + threadIdx.x * N : nullptr;
    float out[N];
    {
        T g0[3], ft[6];
#pragma unroll
        for (int k = 0; k < 3; ++k) g0[k] = (T)a.tip.g0[k];
        if (TIP) {
#pragma unroll
            for (int k = 0; k < 6; ++k) ft[k] = (T)a.tip.ftip[k];
        }
        const T *ftp = TIP ? ft : nullptr;
        T tau[N];
        Store st;
        st.base = wsm + 2 * gl;
#if MPK_FUSED_LAZY_SINCOS
        TrajInLazy<N, T> in{ts, tab, a.jlim, stage};
#else
        {
            T th[N];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                // start + s * delta, one rounding to float32, clipped (bounds are infinite without limits)
                const float pj = clip_f32((float)rn_add(rn_mul(ts.s, tab[2 * j + 1]), tab[2 * j]), a.jlim.lo[j], a.jlim.hi[j]);
                if (WRITE) stage[j] = pj;
                th[j] = (T)pj;
            }
            st.template precompute<REV>(rb, th);
        }
        TrajInPre<N, T> in{ts.sd, ts.sdd, tab, stage};
#endif
        rnea<T, N, GEN, REV, GEO>(rb, in, g0, ftp, tau, st);
#pragma unroll
        for (int j = 0; j < N; ++j) out[j] = clip_f32((float)tau[j], a.tlim.lo[j], a.tlim.hi[j]);
    }
    fused_group_sync();  // every thread of the group is done with its link state: the bytes become the staging tile
#pragma unroll
    for (int j = 0; j < N; ++j) sm[gl * N + j] = out[j];
    fused_group_sync();
    if (live_n > 0) group_tile_store<G>(a.tau + pw * N, sm, live_n * N);
    if (WRITE) {
        __syncthreads();
        const int64_t p0 = (int64_t)blockIdx.x * kDynThreads;
        const int64_t remb = a.P - p0;
        const int cnt = (int)(remb < kDynThreads ? remb : kDynThreads) * N;
        if (a.pos) tile_store(a.pos + p0 * N, traj_sm, cnt);
        if (a.vel) tile_store(a.vel + p0 * N, traj_sm + kDynThreads * N, cnt);
        if (a.acc) tile_store(a.acc + p0 * N, traj_sm + 2 * kDynThreads * N, cnt);
    }
}

// ---- mass matrix -------------------------------------------------------------------
template <int N, bool GEN, bool REV, unsigned GEO = 0>
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
        joint_cs<double, N, REV>(rb, th, q);
        double Mm[N][N];
        mass_matrix<double, N, GEN, REV, GEO>(rb, th, q, Mm);
        double *row = buf + lane * WarpStage<N * N>::S;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) row[i * N + j] = Mm[i][j];
    }
    WarpStage<N * N>::flush(buf, a.out + pw * (N * N), rows);
}

// ---- per-point forward dynamics --------------------------------------------------------
template <int N, bool GEN, bool REV, unsigned GEO = 0>
__global__ void __launch_bounds__(kDynThreads)
    forward_dynamics_kernel(const __grid_constant__ RobotPack<double, N> rb, const FdArgs a) {
    const int64_t p = (int64_t)blockIdx.x * kDynThreads + threadIdx.x;
    if (p >= a.P) return;
    double th[N], dth[N], tau[N], dd[N];
    load_row<N>(a.th, MPK_F64, a.vec, p, th);
    load_row<N>(a.dth, MPK_F64, a.vec, p, dth);
    load_row<N>(a.tau, MPK_F64, a.vec, p, tau);
    double ft[6];
    const double *ftp = load_tip(a.tip, p, ft);
    forward_dynamics<double, N, GEN, REV, 0, GEO>(rb, th, dth, tau, a.tip.g0, ftp, dd);
    store_row_f64<N>(a.out, a.vec, p, dd);
}

// ---- forward-dynamics rollouts, parallel across trajectories, sequential in time ------
// Replaces forward_dynamics_trajectory's CPU loop (planning/trajectory_dynamics.py:580-708):
// per row i >= 1, intRes semi-implicit Euler sub-steps of
//   ddth = M(th)^-1 (taumat[i] - c - g - Js^T Ftipmat[i]);  dth += ddth*dts;  th += dth*dts;
//   th = clip(th, float32 limits);
// rows are stored as float32 and the acceleration row is the last sub-step's.  Row 0 is the
// initial state with zero acceleration and taumat[0] is never used.  The Euler updates use
// explicit round-to-nearest multiplies and adds (no FMA contraction) like the reference's
// NumPy.  One thread owns one trajectory; state, mass matrix and LDL^T factor live in registers
// (a shared-memory hand-over between the recursion, the mass matrix and the solve that halves
// the register count and doubles the resident warps was measured 15-40 % SLOWER: the step is a
// serial dependency chain and the extra shared-memory round trips lengthen it;
// profiles/r1_variants.md).  The torque row of the next step is fetched with cp.async into a
// shared-memory staging column a whole step ahead, so its DRAM latency is off the chain.
template <int N>
__device__ __forceinline__ void store_state(float *o, int64_t row, const double (&x)[N]) {
    float *r = o + row * N;
#pragma unroll
    for (int j = 0; j < N; ++j) r[j] = (float)x[j];
}

// Asynchronous copy of one torque row (N values of 4 or 8 bytes) from global to the lane's
// shared-memory staging column: no registers are held while the row is in flight, and the
// copy is issued a whole step ahead of its use.
template <int N>
__device__ __forceinline__ void tau_row_async(double *stage, const void *taumat, int dtype, int64_t row) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(stage);
    if (dtype == MPK_F64) {
        const double *src = static_cast<const double *>(taumat) + row * N;
#pragma unroll
        for (int j = 0; j < N; ++j)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + j * 32 * 8), "l"(src + j) : "memory");
    } else {
        const float *src = static_cast<const float *>(taumat) + row * N;
#pragma unroll
        for (int j = 0; j < N; ++j)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + j * 32 * 8), "l"(src + j) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int N>
__device__ __forceinline__ void tau_row_take(const double *stage, int dtype, double (&tau)[N]) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const volatile double *p = stage + j * 32;
        tau[j] = dtype == MPK_F64 ? *p : (double)*reinterpret_cast<const volatile float *>(p);
    }
}

// shared memory per warp: two torque-row stages + two words per lane for the loop state (the
// rollout's row base and its step counter)
template <int N>
constexpr size_t rollout_smem_per_warp() {
    return sizeof(double) * 32 * (N * 2 + 2);
}

// TIP: rows of Ftipmat are applied (else every tip-wrench term is compiled out).
template <int N, bool GEN, bool REV, bool TIP, unsigned GEO = 0>
__global__ void __launch_bounds__(kRolloutThreads, MPK_FD_MINBLOCKS)
    fd_rollout_kernel(const __grid_constant__ RobotPack<double, N> rb, const RolloutArgs a) {
    extern __shared__ __align__(16) double fsm[];
    double *stage = fsm + (threadIdx.x >> 5) * (32 * (N * 2 + 2)) + (threadIdx.x & 31);  // + (step & 1) * 32 * N
    // (with block barriers in the step every thread must stay: surplus threads of the last block
    // recompute the last rollout and store nothing)
    const int64_t b_ = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = b_ < a.B;
    if (MPK_FD_PHASES == 0 && !live) return;
    const int64_t b = live ? b_ : a.B - 1;
    double th[N], dth[N], last[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        th[j] = a.th0[b * N + j];
        dth[j] = a.dth0[b * N + j];
        last[j] = 0.0;
    }
    // The step body needs every register; what the loop carries per thread besides the state --
    // the rollout's row base and the step counter -- is parked in shared memory and re-read at
    // the top of each step.  (Left to the compiler, the row pointers derived from the base, and
    // then the 64-bit counter itself, are spilled to LOCAL memory, whose reloads miss L1 behind
    // the torque stream: ncu long_scoreboard 21 % of the samples, all on the reload at the loop
    // head, profiles/r1_variants.md E.)
    volatile int64_t *base_slot = reinterpret_cast<volatile int64_t *>(stage + 32 * N * 2);
    volatile int64_t *step_slot = reinterpret_cast<volatile int64_t *>(stage + 32 * (N * 2 + 1));
    {
        const int64_t base = b * a.N;
        *base_slot = base;
        *step_slot = 1;
        if (MPK_FD_PHASES == 0 || live) {
            store_state<N>(a.pos, base, th);
            store_state<N>(a.vel, base, dth);
            store_state<N>(a.acc, base, last);
        }
        if (a.N > 1) tau_row_async<N>(stage + 32 * N, a.taumat, a.tau_dtype, base + 1);
    }
    for (;;) {
        const int64_t i = *step_slot;
        if (i >= a.N) break;
        *step_slot = i + 1;
        if (MPK_FD_PHASES >= 1) __syncthreads();
        const int64_t base = *base_slot;
        double tau[N];
        tau_row_take<N>(stage + (i & 1) * 32 * N, a.tau_dtype, tau);
        // the torque row of step i + 1 travels while step i is being computed
        if (i + 1 < a.N) tau_row_async<N>(stage + ((i + 1) & 1) * 32 * N, a.taumat, a.tau_dtype, base + i + 1);
        double ft[6];
        if (TIP) {
#pragma unroll
            for (int k = 0; k < 6; ++k) ft[k] = __ldg(a.ftipmat + (base + i) * 6 + k);
        }
        const double *ftp = TIP ? ft : nullptr;
#pragma unroll
        for (int j = 0; j < N; ++j) last[j] = 0.0;
        for (int r = 0; r < a.intRes; ++r) {
            double dd[N];
            forward_dynamics<double, N, GEN, REV, MPK_FD_PHASES, GEO>(rb, th, dth, tau, a.g0, ftp, dd);
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
        const int64_t row = *base_slot + *step_slot - 1;
        if (MPK_FD_PHASES == 0 || live) {
            store_state<N>(a.pos, row, th);
            store_state<N>(a.vel, row, dth);
            store_state<N>(a.acc, row, last);
        }
    }
}
// ---- the same rollouts with each Euler step split across a PAIR of warps ----------------------
// A step is a ~1950-instruction fp64 dependency chain; with 8 warps per SM (255 registers) nothing
// hides its latency.  Here lane l of warp A and lane l of warp B of a 64-thread block own rollout
// 32 blockIdx.x + l together:
//   A: joint sin / cos -> (c, s) to shared memory | bias forces (Newton-Euler, ddtheta = 0) -> shared
//      memory | ... waits for ddtheta | Euler update, limits, row stores
//   B: torque row (cp.async, one step ahead) | waits for (c, s) | mass matrix (CRBA) + LDL^T
//      factorisation | waits for the bias forces | ddtheta = M^-1 (tau - bias) -> shared memory
// so the bias forces and the mass matrix -- the two halves of the chain -- run side by side, and
// each warp holds about half of the state.  Hand-over by named barriers (bar.arrive on the
// producer, bar.sync on the consumer, 64 threads each): 1 = (c, s) ready, 2 = bias ready,
// 3 = ddtheta ready.  Same arithmetic as fd_rollout_kernel (the same bits over short horizons:
// test_rollout_kernels_agree; to the last float32 bit of a few entries in 1e7 over 1000 steps:
// test_rollout_kernels_long_horizon -- ptxas contracts the same expressions differently per kernel).
// Plain revolute chains with
// rigid links and no tip wrench; everything else takes fd_rollout_kernel.
// (no fence: st.shared; bar.arrive | bar.sync; ld.shared is the PTX ISA's own producer / consumer
// pattern, and a MEMBAR here would also wait for the row stores and the cp.async in flight)
__device__ __forceinline__ void pair_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void pair_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

#ifndef MPK_FD_PAIR_MINBLOCKS
#define MPK_FD_PAIR_MINBLOCKS 4
#endif
// MPK_FD_PAIR_REREAD_MAX (tuning knob): instantiations compiled for more resident blocks than this skip A's re-read of (c, s)
// from shared memory (the re-read keeps ptxas from scheduling the bias recursion ahead of the hand-over, which matters when
// B has nothing else to run -- one wave; with 12 warps per SM the schedulers always have another warp, and without the
// re-read 40,000 rollouts take 7.25 instead of 7.94 ms, 65,536 the same 11.2 ms)
// MPK_FD_PAIR_WAVES_MINBLOCKS: resident blocks per SM of the instantiation that takes batches of several waves
#ifndef MPK_FD_PAIR_WAVES_MINBLOCKS
#define MPK_FD_PAIR_WAVES_MINBLOCKS 6
#endif
#ifndef MPK_FD_PAIR_REREAD_MAX
#define MPK_FD_PAIR_REREAD_MAX 4
#endif
constexpr int kRolloutPairBlocksPerSm = MPK_FD_PAIR_MINBLOCKS;

// doubles of shared memory per block: (c, s) 2 N, bias N, ddtheta N, two torque-row stages 2 N
template <int N>
constexpr size_t rollout_pair_smem() {
    return sizeof(double) * 32 * (6 * N);
}

// MINB: resident blocks per SM the kernel is compiled for (4: up to 255 registers, batches of one wave; 6: a
// 168-register cap, 12 warps per SM, for batches of several waves -- 65,536 rollouts 11.2 ms against 12.0 ms
// for the single-warp kernel and 12.3 ms at MINB = 4, profiles/r2_variants.md C).
template <int N, unsigned GEO = 0, int MINB = MPK_FD_PAIR_MINBLOCKS>
__global__ void __launch_bounds__(64, MINB)
    fd_rollout_pair_kernel(const __grid_constant__ RobotPack<double, N> rb, const RolloutArgs a) {
    extern __shared__ __align__(16) double psm[];
    const int lane = threadIdx.x & 31;
    double *cs = psm + lane;               // [2 N][32]
    double *bias_s = cs + 32 * 2 * N;      // [N][32]
    double *dd_s = bias_s + 32 * N;        // [N][32]
    double *stage = dd_s + 32 * N;         // [2][N][32]
    const int64_t b_ = (int64_t)blockIdx.x * 32 + lane;
    const bool live = b_ < a.B;            // surplus lanes of the last block shadow the last rollout
    const int64_t b = live ? b_ : a.B - 1;
    const int64_t base = b * a.N;
    if (threadIdx.x < 32) {
        // ---- warp A: state, joint rotations, bias forces, integration, output rows ----
        double th[N], dth[N], last[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            th[j] = a.th0[b * N + j];
            dth[j] = a.dth0[b * N + j];
            last[j] = 0.0;
        }
        // The chain that bounds a step is  (c, s) [A] -> mass matrix, factorisation, solve [B] -> Euler
        // update [A] -> (c, s) of the next step; everything else has to stay OFF it:
        //  * the row of the step just finished is stored after (c, s) of the next step has been handed
        //    over (th, dth, last are unchanged until the next update), not before;
        //  * the bias recursion takes (c, s) from the shared-memory copy, re-read AFTER bar.arrive: fed
        //    from the registers, ptxas schedules most of the recursion's arithmetic ahead of the barrier
        //    (a bar.arrive orders memory, not register arithmetic) and warp B waits for (c, s) through most
        //    of the recursion -- 47 % of B's stall samples in profiles/r2_ncu_fd_rollout_pair_source.md.
        int64_t pending = base;  // row that th / dth / last still have to be stored to
        for (int64_t i = 1; i < a.N; ++i) {
            for (int r = 0; r < a.intRes; ++r) {
                RegStorePre<double, N> st_;
                joint_cs_all<double, N, true>(rb, th, st_.q);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    cs[32 * j] = st_.q.c[j];
                    cs[32 * (N + j)] = st_.q.s[j];
                }
                pair_arrive(1);
                if (pending >= 0) {
                    if (live) {
                        store_state<N>(a.pos, pending, th);
                        store_state<N>(a.vel, pending, dth);
                        store_state<N>(a.acc, pending, last);
                    }
                    pending = -1;
                }
                if (MINB <= MPK_FD_PAIR_REREAD_MAX) {
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        st_.q.c[j] = *(volatile double *)(cs + 32 * j);
                        st_.q.s[j] = *(volatile double *)(cs + 32 * (N + j));
                    }
                }
                double bias[N];
                ArrayInNoAcc<double, N> in{th, dth};
                rnea<double, N, false, true, GEO>(rb, in, a.g0, nullptr, bias, st_);
#pragma unroll
                for (int j = 0; j < N; ++j) bias_s[32 * j] = bias[j];
                pair_arrive(2);
                pair_wait(3);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const double dd = *(volatile double *)(dd_s + 32 * j);
                    dth[j] = rn_add(dth[j], rn_mul(dd, a.dts));
                    double x = rn_add(th[j], rn_mul(dth[j], a.dts));
                    if (a.lim.on) {
                        const double lo = (double)a.lim.lo[j], hi = (double)a.lim.hi[j];
                        x = x < lo ? lo : (x > hi ? hi : x);
                    }
                    th[j] = x;
                    last[j] = dd;
                }
                if (r + 1 == a.intRes) pending = base + i;
            }
        }
        if (live && pending >= 0) {
            store_state<N>(a.pos, pending, th);
            store_state<N>(a.vel, pending, dth);
            store_state<N>(a.acc, pending, last);
        }
    } else {
        // ---- warp B: torque rows, mass matrix, factorisation, solve ----
        if (a.N > 1) tau_row_async<N>(stage + 32 * N, a.taumat, a.tau_dtype, base + 1);
        for (int64_t i = 1; i < a.N; ++i) {
            double tau[N];
            tau_row_take<N>(stage + (i & 1) * 32 * N, a.tau_dtype, tau);
            if (i + 1 < a.N) tau_row_async<N>(stage + ((i + 1) & 1) * 32 * N, a.taumat, a.tau_dtype, base + i + 1);
            for (int r = 0; r < a.intRes; ++r) {
                JointCS<double, N> q;
                pair_wait(1);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    q.c[j] = *(volatile double *)(cs + 32 * j);
                    q.s[j] = *(volatile double *)(cs + 32 * (N + j));
                    q.d[j] = rb.d[j];
                }
                double Mm[N][N], dinv[N], dd[N];
                crba<double, N, true, GEO>(rb, q, Mm);
                ldlt_factor<double, N, true>(Mm, dinv);
                pair_wait(2);
#pragma unroll
                for (int j = 0; j < N; ++j) dd[j] = tau[j] - *(volatile double *)(bias_s + 32 * j);
                ldlt_apply<double, N>(Mm, dinv, dd);
#pragma unroll
                for (int j = 0; j < N; ++j) dd_s[32 * j] = dd[j];
                pair_arrive(3);
            }
        }
    }
}

// ---- ... and across THREE warps -------------------------------------------------------------
// After the pair kernel's hand-over fix the step's longest path runs through warp A: Euler update,
// seven sin / cos, the row stores, the bias recursion, then B's solve.  A third warp S takes what does
// not have to be there: it keeps its own copy of the state (the same Euler update on the same inputs:
// the same bits), stores the three output rows, and evaluates sin / cos of the OUTER joints -- the ones
// the mass-matrix recursion (which starts at the tip) needs first -- while A evaluates the inner ones,
// which its own recursion (which starts at the base) needs first.
//   A: Euler | sin / cos of joints [0, NA) -> shared | ... (c, s) of S | bias forces -> shared | ... ddtheta
//   S: Euler | sin / cos of joints [NA, N) -> shared | row stores | ... ddtheta
//   B: torque row | ... (c, s) | CRBA + LDL^T | ... bias | solve -> ddtheta to shared
// Barriers: 1 = A's (c, s) (A arrives, B waits: 64), 4 = S's (c, s) (S arrives, A and B wait: 96),
// 2 = bias (A arrives, B waits: 64), 3 = ddtheta (B arrives, A and S wait: 96).
// The pair kernel takes the batches between 2 x 32 x SMs and 4 x 32 x SMs rollouts.
template <int N, int LO, int HI>
__device__ __forceinline__ void joint_cs_part(const RobotPack<double, N> &rb, const double (&th)[N], JointCS<double, N> &q) {
    bool near = true;
#pragma unroll
    for (int i = LO; i < HI; ++i) near = near && sincos_is_near(rb.phi[i] + th[i]);
    if (near) {
        // (one basic block: the independent dependency chains interleave)
#pragma unroll
        for (int i = LO; i < HI; ++i) sincos_near(rb.trig, rb.phi[i] + th[i], &q.s[i], &q.c[i]);
    } else {
#pragma unroll
        for (int i = LO; i < HI; ++i) sincos_pack(rb.trig, rb.phi[i] + th[i], &q.s[i], &q.c[i]);
    }
#pragma unroll
    for (int i = LO; i < HI; ++i) q.d[i] = rb.d[i];
}

template <int N>
__device__ __forceinline__ void euler_update(const RolloutArgs &a, const double *dd_s, double (&th)[N], double (&dth)[N],
                                             double (&last)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double dd = *(volatile const double *)(dd_s + 32 * j);
        dth[j] = rn_add(dth[j], rn_mul(dd, a.dts));
        double x = rn_add(th[j], rn_mul(dth[j], a.dts));
        if (a.lim.on) {
            const double lo = (double)a.lim.lo[j], hi = (double)a.lim.hi[j];
            x = x < lo ? lo : (x > hi ? hi : x);
        }
        th[j] = x;
        last[j] = dd;
    }
}

// Block layout: 8 warp slots, two groups of 32 rollouts, one block per SM.  Warps are dealt to the SM's
// four schedulers round robin, so the slot order decides who shares one.  B0 B1 A0 A1 - - S1 S0 (layout 1):
// the two B warps get a scheduler each, A of one group shares with S of the other, two slots stay empty
// (their warps exit at once).  Measured alternatives (iiwa14, 1000 steps, profiles/r2_variants.md):
// three-warp blocks, two per SM (A B S | A B S: B of the second block lands on A's scheduler): 8,192
// rollouts 2.53 ms; six-warp block B0 B1 A0 A1 S0 S1 (layout 0, S on its own group's B's scheduler):
// 2.36 ms; layout 1: 2.02 ms.  Batches up to 64 x SMs rollouts; `groups` = 1 leaves the second group's
// warps out (batches up to 32 x SMs: one group per SM).
#ifndef MPK_FD_TRIO_LAYOUT
#define MPK_FD_TRIO_LAYOUT 1
#endif
constexpr int kTrioThreads = MPK_FD_TRIO_LAYOUT == 1 ? 256 : 192;
// NA: sin / cos of joints [0, NA) on warp A, of [NA, N) on warp S.  With two groups per SM S shares a scheduler
// with the other group's A and takes the smaller half (NA = (N + 1) / 2); with one group per SM it has a
// scheduler to itself and takes the larger one (NA = N / 2: 2,048 rollouts 1.79 -> 1.62 ms; with two groups that
// split costs 7 %, profiles/r2_variants.md C).
template <int N, unsigned GEO = 0, int NA = (N + 1) / 2>
__global__ void __launch_bounds__(kTrioThreads, 1)
    fd_rollout_trio_kernel(const __grid_constant__ RobotPack<double, N> rb, const RolloutArgs a, const int groups) {
    extern __shared__ __align__(16) double psm[];
    const int lane = threadIdx.x & 31, slot = threadIdx.x >> 5;
    // slot -> (role, group): 0 = A, 1 = B, 2 = S, 3 = none
#if MPK_FD_TRIO_LAYOUT == 1
    const int role = slot < 2 ? 1 : (slot < 4 ? 0 : (slot < 6 ? 3 : 2));
    const int grp = slot < 4 ? (slot & 1) : (slot == 6 ? 1 : 0);
#else
    const int role = slot < 2 ? 1 : (slot < 4 ? 0 : 2);
    const int grp = slot & 1;
#endif
    if (role == 3 || grp >= groups) return;
    double *cs = psm + grp * (32 * 6 * N) + lane;  // [2 N][32]
    double *bias_s = cs + 32 * 2 * N;      // [N][32]
    double *dd_s = bias_s + 32 * N;        // [N][32]
    double *stage = dd_s + 32 * N;         // [2][N][32]
    const int64_t b_ = ((int64_t)blockIdx.x * groups + grp) * 32 + lane;
    const bool live = b_ < a.B;            // surplus lanes shadow the last rollout
    const int64_t b = live ? b_ : a.B - 1;
    const int64_t base = b * a.N;
    const int bar0 = 4 * grp;              // named barriers 1..4 (group 0), 5..8 (group 1)
    const auto pair_arrive = [bar0](int id) { asm volatile("bar.arrive %0, 64;" ::"r"(bar0 + id) : "memory"); };
    const auto pair_wait = [bar0](int id) { asm volatile("bar.sync %0, 64;" ::"r"(bar0 + id) : "memory"); };
    const auto arrive96 = [bar0](int id) { asm volatile("bar.arrive %0, 96;" ::"r"(bar0 + id) : "memory"); };
    const auto wait96 = [bar0](int id) { asm volatile("bar.sync %0, 96;" ::"r"(bar0 + id) : "memory"); };
    if (role == 1) {
        // ---- warp B: torque rows, mass matrix, factorisation, solve ----
        if (a.N > 1) tau_row_async<N>(stage + 32 * N, a.taumat, a.tau_dtype, base + 1);
        for (int64_t i = 1; i < a.N; ++i) {
            double tau[N];
            tau_row_take<N>(stage + (i & 1) * 32 * N, a.tau_dtype, tau);
            if (i + 1 < a.N) tau_row_async<N>(stage + ((i + 1) & 1) * 32 * N, a.taumat, a.tau_dtype, base + i + 1);
            for (int r = 0; r < a.intRes; ++r) {
                JointCS<double, N> q;
                wait96(4);
                pair_wait(1);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    q.c[j] = *(volatile double *)(cs + 32 * j);
                    q.s[j] = *(volatile double *)(cs + 32 * (N + j));
                    q.d[j] = rb.d[j];
                }
                double Mm[N][N], dinv[N], dd[N];
                crba<double, N, true, GEO>(rb, q, Mm);
                ldlt_factor<double, N, true>(Mm, dinv);
                pair_wait(2);
#pragma unroll
                for (int j = 0; j < N; ++j) dd[j] = tau[j] - *(volatile double *)(bias_s + 32 * j);
                ldlt_apply<double, N>(Mm, dinv, dd);
#pragma unroll
                for (int j = 0; j < N; ++j) dd_s[32 * j] = dd[j];
                arrive96(3);
            }
        }
        return;
    }
    double th[N], dth[N], last[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        th[j] = a.th0[b * N + j];
        dth[j] = a.dth0[b * N + j];
        last[j] = 0.0;
    }
    if (role == 0) {
        // ---- warp A: inner joints' sin / cos, bias forces ----
        for (int64_t i = 1; i < a.N; ++i) {
            for (int r = 0; r < a.intRes; ++r) {
                RegStorePre<double, N> st_;
                joint_cs_part<N, 0, NA>(rb, th, st_.q);
#pragma unroll
                for (int j = 0; j < NA; ++j) {
                    cs[32 * j] = st_.q.c[j];
                    cs[32 * (N + j)] = st_.q.s[j];
                }
                pair_arrive(1);
                wait96(4);
                // (all of (c, s) re-read behind the barriers: fed from registers, ptxas schedules the
                // recursion's arithmetic ahead of bar.arrive and B waits for it)
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    st_.q.c[j] = *(volatile double *)(cs + 32 * j);
                    st_.q.s[j] = *(volatile double *)(cs + 32 * (N + j));
                    st_.q.d[j] = rb.d[j];
                }
                double bias[N];
                ArrayInNoAcc<double, N> in{th, dth};
                rnea<double, N, false, true, GEO>(rb, in, a.g0, nullptr, bias, st_);
#pragma unroll
                for (int j = 0; j < N; ++j) bias_s[32 * j] = bias[j];
                pair_arrive(2);
                wait96(3);
                euler_update<N>(a, dd_s, th, dth, last);
            }
        }
    } else {
        // ---- warp S: outer joints' sin / cos, output rows ----
        int64_t pending = base;  // row that th / dth / last still have to be stored to
        for (int64_t i = 1; i < a.N; ++i) {
            for (int r = 0; r < a.intRes; ++r) {
                JointCS<double, N> q;
                joint_cs_part<N, NA, N>(rb, th, q);
#pragma unroll
                for (int j = NA; j < N; ++j) {
                    cs[32 * j] = q.c[j];
                    cs[32 * (N + j)] = q.s[j];
                }
                arrive96(4);
                if (pending >= 0) {
                    if (live) {
                        store_state<N>(a.pos, pending, th);
                        store_state<N>(a.vel, pending, dth);
                        store_state<N>(a.acc, pending, last);
                    }
                    pending = -1;
                }
                wait96(3);
                euler_update<N>(a, dd_s, th, dth, last);
                if (r + 1 == a.intRes) pending = base + i;
            }
        }
        if (live && pending >= 0) {
            store_state<N>(a.pos, pending, th);
            store_state<N>(a.vel, pending, dth);
            store_state<N>(a.acc, pending, last);
        }
    }
}

// ======================================================================================
// launchers per (flavour, joint count, geometry signature)
// ======================================================================================
template <int F, int N, unsigned GEO>
void launch_rnea_n(const mpk_robot *rb, const RneaArgs &a, unsigned grid, cudaStream_t s) {
    constexpr bool GEN = flavour_gen(F), REV = flavour_rev(F);
    if (a.compute_f32) {
        launch_smem_l1(rnea_kernel<float, N, GEN, REV, false, GEO>, grid, kDynThreads, wrench_smem<float, N, GEN, REV>(), 7,
                       s, narrow<N, float>(rb), a);
    } else if (!GEN && !a.dth && !a.ddth && !a.tip.has_ftip) {
        // gravity forces: theta rows only, at-rest recursion (HBM-bound)
        launch_smem_l1(rnea_kernel<double, N, GEN, REV, true, GEO>, grid, kDynThreads, wrench_smem<double, N, GEN, REV>(), 5,
                       s, narrow<N>(rb), a);
    } else {
        launch_smem_l1(rnea_kernel<double, N, GEN, REV, false, GEO>, grid, kDynThreads, wrench_smem<double, N, GEN, REV>(),
                       5, s, narrow<N>(rb), a);
    }
}

template <int F, int N, unsigned GEO>
void launch_traj_rnea_n(const mpk_robot *rb, const TrajRneaArgs &a, unsigned grid, cudaStream_t s) {
    constexpr bool GEN = flavour_gen(F), REV = flavour_rev(F);
    const size_t tab = traj_table_bytes(N, a.B, a.N);
    if (a.compute_f32 && a.tip.has_ftip) {
        launch_smem(traj_rnea_kernel<float, N, GEN, REV, true, false, GEO>, grid, kDynThreads,
                    wrench_smem<float, N, GEN, REV>() + tab, s, narrow<N, float>(rb), a);
    } else if (a.compute_f32) {
        launch_smem(traj_rnea_kernel<float, N, GEN, REV, false, false, GEO>, grid, kDynThreads,
                    wrench_smem<float, N, GEN, REV>() + tab, s, narrow<N, float>(rb), a);
    } else if (a.pos || a.vel || a.acc) {
        // (float64, no tip wrench: the launcher only asks for this variant then)
        launch_smem(traj_rnea_kernel<double, N, GEN, REV, false, true, GEO>, grid, kDynThreads,
                    wrench_smem<double, N, GEN, REV>() + 3 * sizeof(float) * kDynThreads * N + tab, s, narrow<N>(rb), a);
    } else if (a.tip.has_ftip) {
        launch_smem(traj_rnea_kernel<double, N, GEN, REV, true, false, GEO>, grid, kDynThreads,
                    wrench_smem<double, N, GEN, REV>() + tab, s, narrow<N>(rb), a);
    } else {
        launch_smem(traj_rnea_kernel<double, N, GEN, REV, false, false, GEO>, grid, kDynThreads,
                    wrench_smem<double, N, GEN, REV>() + tab, s, narrow<N>(rb), a);
    }
}

template <int F, int N, unsigned GEO>
void launch_mass_n(const mpk_robot *rb, const MassArgs &a, unsigned grid, cudaStream_t s) {
    constexpr bool GEN = flavour_gen(F), REV = flavour_rev(F);
    launch_smem(mass_matrix_kernel<N, GEN, REV, GEO>, grid, kDynThreads,
                sizeof(double) * WarpStage<N * N>::kDoubles * (kDynThreads / 32), s, narrow<N>(rb), a);
}

template <int F, int N, unsigned GEO>
void launch_fd_point_n(const mpk_robot *rb, const FdArgs &a, unsigned grid, cudaStream_t s) {
    constexpr bool GEN = flavour_gen(F), REV = flavour_rev(F);
    forward_dynamics_kernel<N, GEN, REV, GEO><<<grid, kDynThreads, 0, s>>>(narrow<N>(rb), a);
}

inline int sm_count() {
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
template <int F, int N, unsigned GEO>
void launch_rollout_n(const mpk_robot *rb, const RolloutArgs &a, cudaStream_t s) {
    constexpr bool GEN = flavour_gen(F), REV = flavour_rev(F);
    constexpr int threads = kRolloutThreads;
    const unsigned grid = (unsigned)((a.B + threads - 1) / threads);
#if MPK_FD_PAIR
    // A batch that fits the GPU in one wave runs each step split across warps: a lone warp is bound by its
    // own instruction issue (~2000 instructions per step, one fp64 instruction per 2.6 - 3 cycles:
    // scripts/lone_warp_probe.py), and with less than one warp per scheduler the other schedulers idle.
    // Up to 2 x 32 rollouts per SM: three warps per 32 rollouts (fd_rollout_trio_kernel); up to 4 x 32 per
    // SM: two (fd_rollout_pair_kernel).  Up to 8 x 32 per SM -- one wave of the single-warp kernel -- that
    // kernel wins (28,416: 5.09 against 5.12 - 6.24 ms); beyond, the pair kernel compiled for 6 blocks per SM
    // does (12 warps per SM hide more latency than 8: 65,536 rollouts 11.2 against 12.0 ms).
    // MPK_FD_SPLIT (tuning knob): 1 = single-warp kernel only, 2 = no three-warp kernel, 4 = pair kernel (4
    // blocks per SM) whatever the batch.
    if constexpr (!GEN && REV && N >= 2) {
        static const int knob = [] {
            const char *e = std::getenv("MPK_FD_SPLIT");
            return e ? std::atoi(e) : 3;
        }();
        const int64_t wave = 32 * (int64_t)sm_count();
        const unsigned blocks = (unsigned)((a.B + 31) / 32);
        if (!a.ftipmat && knob != 1 && (a.B <= 4 * wave || knob == 4)) {
            // plain revolute chain, rigid links, no tip wrench
            if (knob == 3 && a.B <= 2 * wave) {
                if (a.B <= wave)
                    launch_smem(fd_rollout_trio_kernel<N, GEO, N / 2>, blocks, kTrioThreads, 2 * rollout_pair_smem<N>(), s,
                                narrow<N>(rb), a, 1);
                else
                    launch_smem(fd_rollout_trio_kernel<N, GEO>, (unsigned)((a.B + 63) / 64), kTrioThreads,
                                2 * rollout_pair_smem<N>(), s, narrow<N>(rb), a, 2);
            }
            else
                launch_smem(fd_rollout_pair_kernel<N, GEO>, blocks, 64, rollout_pair_smem<N>(), s, narrow<N>(rb), a);
            return;
        }
        if (!a.ftipmat && knob != 1 && a.B > 8 * wave) {
            launch_smem(fd_rollout_pair_kernel<N, GEO, MPK_FD_PAIR_WAVES_MINBLOCKS>, blocks, 64, rollout_pair_smem<N>(), s,
                        narrow<N>(rb), a);
            return;
        }
    }
#endif
    if (a.ftipmat) {
        launch_smem(fd_rollout_kernel<N, GEN, REV, true, GEO>, grid, threads, rollout_smem_per_warp<N>() * (threads / 32), s,
                    narrow<N>(rb), a);
    } else {
        launch_smem(fd_rollout_kernel<N, GEN, REV, false, GEO>, grid, threads, rollout_smem_per_warp<N>() * (threads / 32), s,
                    narrow<N>(rb), a);
    }
}

// Dispatch of one launcher family over the joint count (general kernels, GEO = 0).
#define MPK_DISPATCH_N(FN, ...)                      \
    switch (rb->n) {                                 \
        case 1: FN<F, 1, 0>(__VA_ARGS__); break;     \
        case 2: FN<F, 2, 0>(__VA_ARGS__); break;     \
        case 3: FN<F, 3, 0>(__VA_ARGS__); break;     \
        case 4: FN<F, 4, 0>(__VA_ARGS__); break;     \
        case 5: FN<F, 5, 0>(__VA_ARGS__); break;     \
        case 6: FN<F, 6, 0>(__VA_ARGS__); break;     \
        case 7: FN<F, 7, 0>(__VA_ARGS__); break;     \
        case 8: FN<F, 8, 0>(__VA_ARGS__); break;     \
        default: break;                              \
    }
// ... and the kernels of the geometry signatures: the instantiations live in the geometry units
#define MPK_GEO_EXTERN_(n_, g_)                                                                                     \
    extern template void launch_rnea_n<0, n_, g_>(const mpk_robot *, const RneaArgs &, unsigned, cudaStream_t);          \
    extern template void launch_traj_rnea_n<0, n_, g_>(const mpk_robot *, const TrajRneaArgs &, unsigned, cudaStream_t); \
    extern template void launch_mass_n<0, n_, g_>(const mpk_robot *, const MassArgs &, unsigned, cudaStream_t);          \
    extern template void launch_fd_point_n<0, n_, g_>(const mpk_robot *, const FdArgs &, unsigned, cudaStream_t);        \
    extern template void launch_rollout_n<0, n_, g_>(const mpk_robot *, const RolloutArgs &, cudaStream_t);
#ifndef MPK_GEO_UNIT
MPK_GEO_LIST(MPK_GEO_EXTERN_)
#endif
#endif  // MPK_FLAVOUR_KERNELS

}  // namespace mpk
