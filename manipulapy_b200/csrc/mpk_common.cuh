// mpk_common.cuh -- host-side plumbing shared by the launcher translation units:
// the opaque robot handle, error reporting, DOF dispatch and row load/store helpers.
#pragma once
#include <cstdio>
#include <string>

#include "../../include/mpk.h"
#include "mpk_device.cuh"

struct mpk_robot {
    int n;
    int rigid;         // every link inertia is a rigid body at its centre of mass
    int all_revolute;  // no prismatic joint
    int plain;         // all revolute and every link plain Denavit-Hartenberg (beta = 0): flavour 0
    int first_revolute;  // joint 0 is revolute (required by the rigid kernel flavours)
    int has_dynamics;
    double F[MPK_MAX_DOF][12];  // home pose of link frame k in space (R row-major 9, p 3): e^{[S_k] th} F_k = F_k Rz(th)
    unsigned geo;      // link geometry classes (mpk_device.cuh kGeo*), 4 bits per link; 0 unless `plain`
    mpk::RobotPack<double, MPK_MAX_DOF> pack;  // host copy, frames 0..n-1 valid
};

// Link-geometry signatures (mpk_device.cuh "link geometry classes") that have their own kernels:
// X(joints, signature).  Each is compiled in its own translation unit (csrc/dyn_geo.cu / fd_geo.cu with
// -DMPK_GEO_N / -DMPK_GEO_SIG; _build.py reads this list); the flavour-0 launchers pick them by
// (rb->n, rb->geo) and fall back to the general kernels (GEO = 0) for any other robot.
#define MPK_GEO_LIST(X)                                   \
    X(6, 0xd52ad0u)  /* UR3 .. UR16e */                   \
    X(7, 0xd555150u) /* KUKA iiwa 7 / 14 */               \
    X(6, 0xd558d0u)  /* Fanuc CRX-5iA .. CRX-30iA */      \
    X(6, 0xdd1890u)  /* Fanuc LR Mate, M-16iB */          \
    X(6, 0xdd1a90u)  /* ABB IRB 2400 */                   \
    X(7, 0xc444440u) /* Kinova Gen3 */


namespace mpk {

void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
int check_launch(const char *what);

// Narrow the max-DOF host pack to the N (and arithmetic type) the kernel is instantiated for.
template <int N, typename T = double>
inline RobotPack<T, N> narrow(const mpk_robot *rb) {
    RobotPack<T, N> o;
    const auto &s = rb->pack;
    for (int i = 0; i < N; ++i) {
        o.a[i] = s.a[i];
        o.ca[i] = s.ca[i];
        o.sa[i] = s.sa[i];
        o.cb[i] = s.cb[i];
        o.sb[i] = s.sb[i];
        o.phi[i] = s.phi[i];
        o.d[i] = s.d[i];
        o.cphi[i] = s.cphi[i];
        o.sphi[i] = s.sphi[i];
        for (int k = 0; k < 3; ++k) {
            o.h[i][k] = s.h[i][k];
            o.com[i][k] = s.com[i][k];
            o.cg[i][k] = s.cg[i][k];
        }
        for (int k = 0; k < 6; ++k) {
            o.I[i][k] = s.I[i][k];
            o.Ic[i][k] = s.Ic[i][k];
        }
        for (int k = 0; k < 21; ++k) o.G[i][k] = s.G[i][k];
        o.sr[i] = s.sr[i];
        o.st[i] = s.st[i];
        o.m[i] = s.m[i];
        o.mg[i] = s.mg[i];
    }
    for (int k = 0; k < 9; ++k) {
        o.Rb[k] = s.Rb[k];
        o.Ree[k] = s.Ree[k];
    }
    for (int k = 0; k < 3; ++k) {
        o.pb[k] = s.pb[k];
        o.pee[k] = s.pee[k];
    }
    for (int k = 0; k < 17; ++k) o.trig[k] = s.trig[k];
    return o;
}

#define MPK_DISPATCH_DOF(n, ...)                          \
    switch (n) {                                          \
        case 1: { constexpr int N_ = 1; __VA_ARGS__; } break;    \
        case 2: { constexpr int N_ = 2; __VA_ARGS__; } break;    \
        case 3: { constexpr int N_ = 3; __VA_ARGS__; } break;    \
        case 4: { constexpr int N_ = 4; __VA_ARGS__; } break;    \
        case 5: { constexpr int N_ = 5; __VA_ARGS__; } break;    \
        case 6: { constexpr int N_ = 6; __VA_ARGS__; } break;    \
        case 7: { constexpr int N_ = 7; __VA_ARGS__; } break;    \
        case 8: { constexpr int N_ = 8; __VA_ARGS__; } break;    \
        default: return mpk::fail(MPK_EUNSUPPORTED, "dof must be in 1..8"); \
    }

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- per-thread row I/O -------------------------------------------------------
// Row p of a dense (P, N) array of float64 / float32 -> N doubles in registers.
// Each thread owns one contiguous row; rows of consecutive threads are adjacent, so a
// warp touches one contiguous span and every fetched sector is fully used.  16-byte
// vector loads are used when the row size allows (the launcher checks base alignment).
template <int N>
__device__ __forceinline__ void load_row(const void *base, int dtype, bool vec, int64_t p,
                                         double (&out)[N]) {
    if (dtype == MPK_F64) {
        const double *r = static_cast<const double *>(base) + p * N;
        if ((N % 2 == 0) && vec) {
            const double2 *r2 = reinterpret_cast<const double2 *>(r);
#pragma unroll
            for (int k = 0; k < N / 2; ++k) {
                const double2 v = __ldg(r2 + k);
                out[2 * k] = v.x;
                out[2 * k + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < N; ++k) out[k] = __ldg(r + k);
        }
    } else {
        const float *r = static_cast<const float *>(base) + p * N;
        if ((N % 4 == 0) && vec) {
            const float4 *r4 = reinterpret_cast<const float4 *>(r);
#pragma unroll
            for (int k = 0; k < N / 4; ++k) {
                const float4 v = __ldg(r4 + k);
                out[4 * k] = v.x;
                out[4 * k + 1] = v.y;
                out[4 * k + 2] = v.z;
                out[4 * k + 3] = v.w;
            }
        } else if ((N % 2 == 0) && vec) {
            const float2 *r2 = reinterpret_cast<const float2 *>(r);
#pragma unroll
            for (int k = 0; k < N / 2; ++k) {
                const float2 v = __ldg(r2 + k);
                out[2 * k] = v.x;
                out[2 * k + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < N; ++k) out[k] = __ldg(r + k);
        }
    }
}

template <int K>
__device__ __forceinline__ void store_row_f64(double *base, bool vec, int64_t p,
                                              const double (&v)[K]) {
    double *r = base + p * K;
    if ((K % 2 == 0) && vec) {
        double2 *r2 = reinterpret_cast<double2 *>(r);
#pragma unroll
        for (int k = 0; k < K / 2; ++k) r2[k] = make_double2(v[2 * k], v[2 * k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) r[k] = v[k];
    }
}

template <int K>
__device__ __forceinline__ void store_row_f32(float *base, bool vec, int64_t p,
                                              const float (&v)[K]) {
    float *r = base + p * K;
    if ((K % 4 == 0) && vec) {
        float4 *r4 = reinterpret_cast<float4 *>(r);
#pragma unroll
        for (int k = 0; k < K / 4; ++k)
            r4[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    } else if ((K % 2 == 0) && vec) {
        float2 *r2 = reinterpret_cast<float2 *>(r);
#pragma unroll
        for (int k = 0; k < K / 2; ++k) r2[k] = make_float2(v[2 * k], v[2 * k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) r[k] = v[k];
    }
}

struct Limits {
    float lo[MPK_MAX_DOF], hi[MPK_MAX_DOF];
    int on;
};

inline Limits make_limits(const float *lim, int n) {
    Limits L;
    L.on = lim != nullptr;
    for (int j = 0; j < MPK_MAX_DOF; ++j) {
        // (without limits the bounds are infinite, so that code which clips unconditionally is a no-op)
        L.lo[j] = (lim && j < n) ? lim[2 * j] : -INFINITY;
        L.hi[j] = (lim && j < n) ? lim[2 * j + 1] : INFINITY;
    }
    return L;
}

// Time scaling of step t: from the (3, N) table written by time_scaling_table_kernel when the
// caller supplied a workspace, else evaluated in place.  Same arithmetic either way.
__device__ __forceinline__ TimeScale time_scaling_at(const double *table, int64_t t, int64_t N,
                                                     double Tf, int method) {
    if (table) return TimeScale{__ldg(table + t), __ldg(table + N + t), __ldg(table + 2 * N + t)};
    return time_scaling(t, N, Tf, method);
}

// Launches the trajectory kernel (traj.cu) with an already prepared time-scaling table.
int launch_joint_trajectory(int n, int64_t B, int64_t N, const double *start, const double *end,
                            int inputs_f32, double Tf, int method, const float *limits, float *pos,
                            float *vel, float *acc, const double *ts_table, cudaStream_t s);

// Fills the workspace if it is worth it (more than one trajectory); returns the table to use.
const double *prepare_time_scaling(double *scratch, int64_t B, int64_t N, double Tf, int method,
                                   cudaStream_t s);

// Division of a flattened point index by the trajectory length without a 64-bit divide:
// for n < 2^31, floor(n / d) = (n * M) >> s with M = ceil(2^s / d), s = 31 + ceil(log2 d)
// (Granlund-Montgomery); larger batches fall back to the hardware-emulated 64-bit division.
struct FastDiv {
    uint32_t d, M;
    int shift;  // s - 32, or log2(d) when pow2
    int pow2, wide;
};

inline FastDiv make_fastdiv(int64_t d, int64_t max_n) {
    FastDiv f;
    f.d = (uint32_t)d;
    f.M = 0;
    f.shift = 0;
    f.pow2 = 0;
    f.wide = (max_n >= (1LL << 31)) || (d >= (1LL << 31));
    if (f.wide) return f;
    int l = 0;
    while ((1LL << l) < d) ++l;
    if ((1LL << l) == d) {
        f.pow2 = 1;
        f.shift = l;
        return f;
    }
    const int s = 31 + l;
    const unsigned __int128 num = (unsigned __int128)1 << s;
    f.M = (uint32_t)((num + (unsigned __int128)d - 1) / (unsigned __int128)d);
    f.shift = s - 32;
    return f;
}

// (trajectory, step) of flattened point p = b * N + t.
__device__ __forceinline__ void point_coords(const FastDiv &f, int64_t N, int64_t p, int64_t &b,
                                             int64_t &t) {
    if (f.wide) {
        b = p / N;
        t = p - b * N;
    } else {
        const uint32_t n = (uint32_t)p;
        const uint32_t q = f.pow2 ? (n >> f.shift) : (__umulhi(n, f.M) >> f.shift);
        b = q;
        t = n - q * f.d;
    }
}

// Per-warp output staging: each lane owns one row of K values; the warp then writes the
// (up to) 32 rows, which are contiguous in global memory, with fully coalesced stores.  The
// row stride is odd (in words of the element type) so neither phase has shared-memory bank
// conflicts.
template <int K, typename E = double>
struct WarpStage {
    static constexpr int S = (K % 2 == 0) ? K + 1 : K;
    static constexpr int kDoubles = 32 * S;  // elements per warp (the name predates the float variant)
    // gout: global address of the warp's first row; rows: live rows of this warp (<= 32)
    __device__ static __forceinline__ void flush(const E *buf, E *gout, int rows) {
        __syncwarp();
        const int lane = threadIdx.x & 31;
        const int cnt = rows * K;
        for (int e = lane; e < cnt; e += 32) {
            const int r = e / K, k = e - r * K;
            __stcs(gout + e, buf[r * S + k]);
        }
        __syncwarp();
    }
};

// start and (end - start) of joint j of trajectory b, in the precision the reference uses.
__device__ __forceinline__ void endpoint(const double *start, const double *end, int inputs_f32,
                                         int64_t idx, double &st, double &dth) {
    const double s = __ldg(start + idx), e = __ldg(end + idx);
    if (inputs_f32) {
        const float s32 = (float)s, e32 = (float)e;
        st = (double)s32;
        dth = (double)rn_fsub(e32, s32);
    } else {
        st = s;
        dth = rn_sub(e, s);
    }
}

// Coalesced copy-out of a block's staged rows: cnt floats from shared memory (16-byte aligned) to o.
// The vector width follows the alignment of the destination (block-uniform): 16-byte vectors for the
// usual freshly allocated result, 8-byte or scalar stores when the caller passed a row-offset view of
// a larger buffer (e.g. its shard of another GPU's peer-mapped result, sharding.PeerRows) whose
// first byte is only 8- or 4-byte aligned.
__device__ __forceinline__ void tile_store(float *o, const float *sm, int cnt) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(o);
    if ((addr & 15u) == 0) {
        const int n4 = cnt >> 2;
        const float4 *s4 = reinterpret_cast<const float4 *>(sm);
        float4 *o4 = reinterpret_cast<float4 *>(o);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) __stcs(o4 + i, s4[i]);
        for (int i = (n4 << 2) + threadIdx.x; i < cnt; i += blockDim.x) o[i] = sm[i];
    } else if ((addr & 7u) == 0) {
        const int n2 = cnt >> 1;
        const float2 *s2 = reinterpret_cast<const float2 *>(sm);
        float2 *o2 = reinterpret_cast<float2 *>(o);
        for (int i = threadIdx.x; i < n2; i += blockDim.x) __stcs(o2 + i, s2[i]);
        for (int i = (n2 << 1) + threadIdx.x; i < cnt; i += blockDim.x) o[i] = sm[i];
    } else {
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) __stcs(o + i, sm[i]);
    }
}

// The same for the rows of a group of G threads (a warp or the block; cnt floats from the group's own
// shared-memory slice, 16-byte aligned); only the calling group takes part.
template <int G>
__device__ __forceinline__ void group_tile_store(float *o, const float *sm, int cnt) {
    const int gl = threadIdx.x % G;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(o);
    if ((addr & 15u) == 0) {
        const int n4 = cnt >> 2;
        const float4 *s4 = reinterpret_cast<const float4 *>(sm);
        float4 *o4 = reinterpret_cast<float4 *>(o);
#pragma unroll 2
        for (int i = gl; i < n4; i += G) __stcs(o4 + i, s4[i]);
        for (int i = (n4 << 2) + gl; i < cnt; i += G) o[i] = sm[i];
    } else if ((addr & 7u) == 0) {
        const int n2 = cnt >> 1;
        const float2 *s2 = reinterpret_cast<const float2 *>(sm);
        float2 *o2 = reinterpret_cast<float2 *>(o);
#pragma unroll 2
        for (int i = gl; i < n2; i += G) __stcs(o2 + i, s2[i]);
        for (int i = (n2 << 1) + gl; i < cnt; i += G) o[i] = sm[i];
    } else {
#pragma unroll 2
        for (int i = gl; i < cnt; i += G) __stcs(o + i, sm[i]);
    }
}

}  // namespace mpk
