// traj.cu -- joint-trajectory generation (cubic / quintic time scaling).
//
// Replaces the reference's Numba CPU kernel (planning/trajectory.py:15-75), its clip
// (:311-313), and all seven Numba CUDA trajectory kernels
// (cuda_kernels/trajectory_kernels.py:154-518, 763-831).  HBM-write-bound: 12 n bytes
// per point out, (B, n) endpoints in.  One thread owns one time step of one trajectory
// for all joints, so the three divisions of the time scaling are paid once per n
// outputs; a block covers 128 consecutive points of the flattened (B*N) point index
// and stages its rows through shared memory so that every global store is a
// fully-used 16-byte-per-lane coalesced transaction (tile offsets are multiples of
// 128 n floats, hence always 16-byte aligned).
#include "mpk_common.cuh"

namespace mpk {

constexpr int kTrajThreads = 256;
// Store flavour of the row stream (tuning knob).  In the bare store micro-benchmark plain write-back
// stores reach 6.6 TB/s and streaming (.cs) stores 6.4-6.5 TB/s (profiles/r2_peaks_store_and_fp64_patterns.json);
// in this kernel the streaming flavour is the faster one by 1-2 % (0.1486 against 0.1505 ms).
#ifndef MPK_TRAJ_STREAMING_STORES
#define MPK_TRAJ_STREAMING_STORES 1
#endif
#if MPK_TRAJ_STREAMING_STORES
#define TRAJ_STORE(p, v) __stcs(p, v)
#else
#define TRAJ_STORE(p, v) (*(p) = (v))
#endif

struct TrajArgs {
    int64_t B, N, P;
    FastDiv div;
    const double *start, *end;
    int inputs_f32;
    double Tf;
    int method;
    Limits lim;
    float *pos, *vel, *acc;
    const double *ts_table;
    int table_rows;  // trajectories one warp's 32 consecutive points can touch: min(B, 31 / N + 2)
};

__global__ void time_scaling_table_kernel(int64_t N, double Tf, int method, double *table) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    const TimeScale ts = time_scaling(t, N, Tf, method);
    table[t] = ts.s;
    table[N + t] = ts.sd;
    table[2 * N + t] = ts.sdd;
}

const double *prepare_time_scaling(double *scratch, int64_t B, int64_t N, double Tf, int method,
                                   cudaStream_t s) {
    if (!scratch || B < 2 || N < 1) return nullptr;
    time_scaling_table_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(N, Tf, method, scratch);
    return scratch;
}

// One thread = one (trajectory, step) point, all joints.  Each warp stages its 32 rows of the
// three outputs in its own shared-memory slice and writes them out with 16-byte-per-lane
// coalesced stores (a warp's rows are contiguous: 32 N floats, a multiple of 128 B).
// Only __syncwarp: warps never wait for each other.
//
// The kernel is a pure 12 N bytes-per-point write stream, and what stood between it and the write
// bandwidth of HBM (6.6 TB/s with plain 16-byte stores, profiles/r2_peaks_store_and_fp64_patterns.json)
// was the conversion unit: 3 N double -> float conversions per point are inherent (one rounding to
// float32 per output), but another 3 N went into re-deriving the endpoints (float32 rounding of start /
// end, their difference, back to double) in EVERY thread -- at 16 conversions per clock per SM that is
// more XU time than the stores take (ncu round 1: XU 58 % busy at 64 % of the copy bandwidth).  A warp's
// 32 consecutive points belong to at most 31 / N + 2 trajectories, so each warp now derives those
// endpoints once into a small shared-memory table and its lanes read (start, delta) pairs from there.
__host__ __device__ inline int traj_warp_table_rows(int64_t B, int64_t N) {
    int64_t k = 31 / (N > 0 ? N : 1) + 2;
    return (int)(k > B ? B : k);
}

template <int N>
__global__ void __launch_bounds__(kTrajThreads) traj_kernel(const TrajArgs a) {
    extern __shared__ __align__(16) unsigned char traj_smem[];
    constexpr int kWarps = kTrajThreads / 32;
    float(*sm)[3][32 * N] = reinterpret_cast<float(*)[3][32 * N]>(traj_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *table = reinterpret_cast<double *>(traj_smem + sizeof(float) * kWarps * 3 * 32 * N) + warp * a.table_rows * 2 * N;
    const int64_t pw = (int64_t)blockIdx.x * kTrajThreads + warp * 32;  // warp's first point
    if (pw >= a.P) return;                                               // whole warp out of range
    const int64_t rem = a.P - pw;
    const int live = (int)(rem < 32 ? rem : 32);
    int64_t b, t;
    point_coords(a.div, a.N, lane < live ? pw + lane : pw, b, t);
    // first / last trajectory of the warp's points: lanes 0 and live - 1 already know them
    const int64_t b0 = __shfl_sync(0xffffffffu, b, 0), b1 = __shfl_sync(0xffffffffu, b, live - 1);
    {
        const int entries = (int)(b1 - b0 + 1) * N;
        for (int e = lane; e < entries; e += 32) {
            double st, dth;
            endpoint(a.start, a.end, a.inputs_f32, b0 * N + e, st, dth);
            table[2 * e] = st;
            table[2 * e + 1] = dth;
        }
    }
    __syncwarp();
    if (lane < live) {
        const TimeScale ts = time_scaling_at(a.ts_table, t, a.N, a.Tf, a.method);
        const double *tab = table + (b - b0) * (2 * N);
        float pr[N], vr[N], ar[N];
#pragma unroll
        for (int j = 0; j < N; ++j)
            traj_point(ts, tab[2 * j], tab[2 * j + 1], a.lim.lo[j], a.lim.hi[j], true, pr[j], vr[j], ar[j]);
        // row stores: 8-byte vectors when N is even (conflict-free at a 4 N byte lane stride)
        float *r0 = &sm[warp][0][lane * N], *r1 = &sm[warp][1][lane * N], *r2 = &sm[warp][2][lane * N];
        if (N % 2 == 0) {
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                *reinterpret_cast<float2 *>(r0 + j) = make_float2(pr[j], pr[j + 1]);
                *reinterpret_cast<float2 *>(r1 + j) = make_float2(vr[j], vr[j + 1]);
                *reinterpret_cast<float2 *>(r2 + j) = make_float2(ar[j], ar[j + 1]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                r0[j] = pr[j];
                r1[j] = vr[j];
                r2[j] = ar[j];
            }
        }
    }
    __syncwarp();
    const int64_t off = pw * N;
    if (live == 32) {
        // full warp (all but the batch's last one): 8 N 16-byte vectors per output, trip count known at compile time
        constexpr int kVec = 8 * N;
        auto flush = [&](float *out, int k) {
            if (!out) return;
            const float4 *s4 = reinterpret_cast<const float4 *>(sm[warp][k]);
            float4 *o4 = reinterpret_cast<float4 *>(out + off);
#pragma unroll
            for (int i0 = 0; i0 < kVec; i0 += 32) {
                if (i0 + 32 <= kVec || lane < kVec - i0) TRAJ_STORE(o4 + i0 + lane, s4[i0 + lane]);
            }
        };
        flush(a.pos, 0);
        flush(a.vel, 1);
        flush(a.acc, 2);
        return;
    }
    const int cnt = live * N;  // floats of this warp per output
    auto flush_tail = [&](float *out, int k) {
        if (!out) return;
        float *o = out + off;
        const float4 *s4 = reinterpret_cast<const float4 *>(sm[warp][k]);
        float4 *o4 = reinterpret_cast<float4 *>(o);
        const int n4 = cnt >> 2;
#pragma unroll 1
        for (int i = lane; i < n4; i += 32) TRAJ_STORE(o4 + i, s4[i]);
#pragma unroll 1
        for (int i = (n4 << 2) + lane; i < cnt; i += 32) o[i] = sm[warp][k][i];
    };
    flush_tail(a.pos, 0);
    flush_tail(a.vel, 1);
    flush_tail(a.acc, 2);
}

int launch_joint_trajectory(int n, int64_t B, int64_t N, const double *start, const double *end,
                            int inputs_f32, double Tf, int method, const float *limits, float *pos,
                            float *vel, float *acc, const double *ts_table, cudaStream_t s) {
    for (float *o : {pos, vel, acc})
        if (o && !aligned16(o)) return fail(MPK_EINVAL, "outputs must be 16-byte aligned");
    TrajArgs a;
    a.B = B;
    a.N = N;
    a.P = B * N;
    a.div = make_fastdiv(N, a.P);
    a.start = start;
    a.end = end;
    a.inputs_f32 = inputs_f32;
    a.Tf = Tf;
    a.method = method;
    a.lim = make_limits(limits, n);
    a.pos = pos;
    a.vel = vel;
    a.acc = acc;
    a.ts_table = ts_table;
    a.table_rows = traj_warp_table_rows(B, N);
    const int64_t blocks = (a.P + kTrajThreads - 1) / kTrajThreads;
    if (blocks > 0x7fffffffLL) return fail(MPK_EINVAL, "B*N exceeds the grid limit (2^39 points)");
    MPK_DISPATCH_DOF(n, {
        const size_t smem = sizeof(float) * (kTrajThreads / 32) * 3 * 32 * N_ +
                            sizeof(double) * (kTrajThreads / 32) * traj_warp_table_rows(B, N) * 2 * N_;
        auto kern = traj_kernel<N_>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<(unsigned)blocks, kTrajThreads, smem, s>>>(a);
    });
    return check_launch("joint_trajectory");
}

// ---- Cartesian straight-line trajectories ------------------------------------------------
struct CartArgs {
    int64_t B, N, P;
    FastDiv div;
    const double *Xs, *Xe;
    double Tf;
    int method;
    float *pos, *vel, *acc, *orient;
};

// One thread = one (trajectory, step): 3 + 3 + 3 + 9 float32 out, staged per warp and flushed
// with coalesced stores like the joint-space kernel.
__global__ void __launch_bounds__(kTrajThreads) cartesian_kernel(const CartArgs a) {
    __shared__ __align__(16) float sm[kTrajThreads / 32][32 * 9];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t pw = (int64_t)blockIdx.x * kTrajThreads + warp * 32;
    if (pw >= a.P) return;
    const int64_t p = pw + lane;
    float pos[3], vel[3], acc[3], R[9];
    if (p < a.P) {
        int64_t b, t;
        point_coords(a.div, a.N, p, b, t);
        cartesian_point(a.Xs + 16 * b, a.Xe + 16 * b, t, a.N, a.Tf, a.method, pos, vel, acc, R);
    }
    const int64_t rem = a.P - pw;
    const int rows = (int)(rem < 32 ? rem : 32);
    float *buf = sm[warp];
    const float *src3[3] = {pos, vel, acc};
    float *dst3[3] = {a.pos, a.vel, a.acc};
    // A full warp's rows are 384 B (1152 B of orientations) starting on a multiple of that: 16-byte vector
    // stores when the array itself is 16-byte aligned, scalar stores for the ragged last warp.
    const auto flush = [&](float *dst, int per_row) {
        float *o = dst + pw * per_row;
        const int cnt = rows * per_row;
        if (rows == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            const float4 *s4 = reinterpret_cast<const float4 *>(buf);
            float4 *o4 = reinterpret_cast<float4 *>(o);
            for (int e = lane; e < cnt / 4; e += 32) __stcs(o4 + e, s4[e]);
        } else {
            for (int e = lane; e < cnt; e += 32) __stcs(o + e, buf[e]);
        }
    };
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (!dst3[k]) continue;
        if (lane < rows) {
#pragma unroll
            for (int j = 0; j < 3; ++j) buf[lane * 3 + j] = src3[k][j];
        }
        __syncwarp();
        flush(dst3[k], 3);
        __syncwarp();
    }
    if (a.orient) {
        if (lane < rows) {
#pragma unroll
            for (int j = 0; j < 9; ++j) buf[lane * 9 + j] = R[j];
        }
        __syncwarp();
        flush(a.orient, 9);
    }
}

}  // namespace mpk

using namespace mpk;

extern "C" int mpk_cartesian_trajectory(int64_t B, int64_t N, const double *Xstart, const double *Xend,
                                        double Tf, int method, float *pos, float *vel, float *acc,
                                        float *orientations, void *stream) {
    if (B < 0 || N < 0) return fail(MPK_EINVAL, "negative size");
    if (B == 0 || N == 0) return MPK_OK;
    if (!Xstart || !Xend) return fail(MPK_EINVAL, "Xstart / Xend are NULL");
    CartArgs a;
    a.B = B;
    a.N = N;
    a.P = B * N;
    a.div = make_fastdiv(N, a.P);
    a.Xs = Xstart;
    a.Xe = Xend;
    a.Tf = Tf;
    a.method = method;
    a.pos = pos;
    a.vel = vel;
    a.acc = acc;
    a.orient = orientations;
    const int64_t blocks = (a.P + kTrajThreads - 1) / kTrajThreads;
    if (blocks > 0x7fffffffLL) return fail(MPK_EINVAL, "B*N exceeds the grid limit");
    cartesian_kernel<<<(unsigned)blocks, kTrajThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("cartesian_trajectory");
}

extern "C" int mpk_joint_trajectory(int n, int64_t B, int64_t N, const double *start,
                                    const double *end, int inputs_f32, double Tf, int method,
                                    const float *limits, float *pos, float *vel, float *acc,
                                    double *ts_scratch, void *stream) {
    if (n < 1 || n > MPK_MAX_DOF) return fail(MPK_EUNSUPPORTED, "dof must be in 1..8");
    if (B < 0 || N < 0) return fail(MPK_EINVAL, "negative size");
    if (B == 0 || N == 0) return MPK_OK;
    if (!start || !end) return fail(MPK_EINVAL, "start/end are NULL");
    if ((N + 255) / 256 > 0x7fffffffLL) return fail(MPK_EINVAL, "N exceeds the grid limit");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const double *table = prepare_time_scaling(ts_scratch, B, N, Tf, method, s);
    return launch_joint_trajectory(n, B, N, start, end, inputs_f32, Tf, method, limits, pos, vel, acc, table, s);
}
