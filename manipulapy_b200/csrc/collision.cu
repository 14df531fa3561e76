// collision.cu -- the collision / limit post-processing hook of joint_trajectory,
// batched: every link's pose, the self-collision test of the reference's CollisionChecker and the
// potential-field nudge of colliding trajectory rows.
//
// Replaces, per trajectory row, the host loop of planning/collision_host.py:40-88:
//   URDF.link_fk over the whole link tree            (urdf/core.py:532-575; batched :577-633)
//   CollisionChecker.check_collision                  (potential_field/collision.py:162-195): for every pair
//       of link hulls not in the allowed-collision set, transform the hull points to the world and test
//       the axis-aligned boxes of the two point sets for overlap (:197-221)
//   PotentialField.compute_gradient with no obstacles (potential_field/fields.py:112-170): attractive term
//   row <- row - 0.01 * gradient in float32, up to 100 times, until the row is collision free.
//
// A link l hangs on actuated joint k(l) (or on the base, k = -1); with W_k the world pose of joint
// frame k AFTER joint k's rotation (the same chain fk_jacobian walks), its pose is W_k C_l with the
// constant C_l = F_k^-1 T_l(0).  The host packs C_l, and the hull points already moved into joint-frame
// coordinates (u = C_l v), so a row costs one chain walk, 9 FMAs + 6 min / max per hull point and the pair
// tests.  One thread owns one row; the per-hull boxes live in a shared-memory column per thread.
// fp64 throughout (the reference's checker is NumPy float64); the nudge itself is float32 like the
// reference's rows.
#include <cstring>
#include <vector>

#include "mpk_common.cuh"

namespace mpk {

constexpr int kColThreads = 64;
constexpr uint32_t kColMagic = 0x6d706b43u;  // "mpkC"

// Packed model (host-built, caller-uploaded; offsets in bytes from the start, all 8-byte aligned)
struct ColHeader {
    uint32_t magic, n, L, H;        // joints, links, hulls
    uint32_t npairs, npoints, pad0, pad1;
    uint64_t off_link_joint;        // int32[L]
    uint64_t off_link_C;            // double[L][12]
    uint64_t off_hull_begin;        // int32[n + 2]: hulls of joint k are [hull_begin[k + 1], hull_begin[k + 2]) (k = -1 first)
    uint64_t off_hull_pts;          // int32[H + 1]: point range of hull h
    uint64_t off_pairs;             // uint8[npairs][2]
    uint64_t off_points;            // double[npoints][3]
    uint64_t bytes;
};

struct ColArgs {
    int64_t P;
    const unsigned char *model;
    const void *theta;
    int theta_dtype;
    // link_fk
    double *T;
    // flags
    uint8_t *flags;
    // avoidance
    float *rows;          // (P, n) float32, in place
    const float *goal;    // (G, n) float32; row p uses goal[p / rows_per_goal]
    int64_t rows_per_goal;
    float gain, step;
    int max_iter;
    int32_t *iters;
};

// World poses of the joint frames, one after the other: after step(i) (X, Y, Z, p) is W_i.
template <int N>
struct Chain {
    double X[3], Y[3], Z[3], p[3];
    __device__ __forceinline__ void start(const RobotPack<double, N> &rb) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            X[r] = rb.Rb[3 * r];
            Y[r] = rb.Rb[3 * r + 1];
            Z[r] = rb.Rb[3 * r + 2];
            p[r] = rb.pb[r];
        }
    }
    __device__ __forceinline__ void step(const RobotPack<double, N> &rb, int i, double c, double s, double dz) {
        if (i > 0) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                p[r] += rb.a[i] * X[r];
                rot_t(rb.ca[i], rb.sa[i], Y[r], Z[r]);
            }
            if (rb.sb[i] != 0.0) {
#pragma unroll
                for (int r = 0; r < 3; ++r) rot_t(rb.cb[i], rb.sb[i], Z[r], X[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            rot_t(c, s, X[r], Y[r]);
            p[r] += dz * Z[r];
        }
    }
};

// Boxes of the hulls of joint k (k = -1: fixed to the base, identity pose) into the thread's column.
template <int N>
__device__ __forceinline__ void hull_boxes(const ColHeader &h, const unsigned char *m, int k, const Chain<N> *ch,
                                           double *box /* [H][6] stride blockDim */) {
    const int *hb = reinterpret_cast<const int *>(m + h.off_hull_begin);
    const int *hp = reinterpret_cast<const int *>(m + h.off_hull_pts);
    const double *pts = reinterpret_cast<const double *>(m + h.off_points);
    const int stride = blockDim.x;
    for (int q = __ldg(hb + k + 1); q < __ldg(hb + k + 2); ++q) {
        double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int e = __ldg(hp + q); e < __ldg(hp + q + 1); ++e) {
            const double u0 = __ldg(pts + 3 * e), u1 = __ldg(pts + 3 * e + 1), u2 = __ldg(pts + 3 * e + 2);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                // R u + p with the reference's summation order (NumPy matmul row, then the translation)
                const double w = ch ? ((ch->X[r] * u0 + ch->Y[r] * u1) + ch->Z[r] * u2) + ch->p[r] : (r == 0 ? u0 : r == 1 ? u1 : u2);
                lo[r] = fmin(lo[r], w);
                hi[r] = fmax(hi[r], w);
            }
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            box[(q * 6 + r) * stride] = lo[r];
            box[(q * 6 + 3 + r) * stride] = hi[r];
        }
    }
}

// CollisionChecker.check_collision of one configuration.
template <int N>
__device__ __forceinline__ bool collides(const RobotPack<double, N> &rb, const ColHeader &h, const unsigned char *m,
                                         const double (&th)[N], double *box) {
    JointCS<double, N> q;
    joint_cs<double, N, false>(rb, th, q);
    hull_boxes<N>(h, m, -1, nullptr, box);
    Chain<N> ch;
    ch.start(rb);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        ch.step(rb, i, q.c[i], q.s[i], q.d[i]);
        hull_boxes<N>(h, m, i, &ch, box);
    }
    const unsigned char *pairs = m + h.off_pairs;
    const int stride = blockDim.x;
    bool hit = false;
    for (unsigned e = 0; e < h.npairs; ++e) {
        const int a = pairs[2 * e], b = pairs[2 * e + 1];
        bool ov = true;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            // max_a >= min_b and max_b >= min_a (collision.py:219-221)
            ov = ov && box[(a * 6 + 3 + r) * stride] >= box[(b * 6 + r) * stride] &&
                 box[(b * 6 + 3 + r) * stride] >= box[(a * 6 + r) * stride];
        }
        hit = hit || ov;
    }
    return hit;
}

template <int N>
__global__ void __launch_bounds__(kColThreads) self_collision_kernel(const __grid_constant__ RobotPack<double, N> rb,
                                                                     const ColArgs a) {
    extern __shared__ __align__(16) double box_sm[];
    const int64_t p = (int64_t)blockIdx.x * kColThreads + threadIdx.x;
    if (p >= a.P) return;
    const ColHeader h = *reinterpret_cast<const ColHeader *>(a.model);
    double th[N];
    load_row<N>(a.theta, a.theta_dtype, false, p, th);
    a.flags[p] = collides<N>(rb, h, a.model, th, box_sm + threadIdx.x) ? 1 : 0;
}

// _apply_collision_avoidance_cpu (planning/collision_host.py:40-88) of one float32 row, in place.
template <int N>
__global__ void __launch_bounds__(kColThreads) collision_avoidance_kernel(const __grid_constant__ RobotPack<double, N> rb,
                                                                          const ColArgs a) {
    extern __shared__ __align__(16) double box_sm[];
    const int64_t p = (int64_t)blockIdx.x * kColThreads + threadIdx.x;
    if (p >= a.P) return;
    const ColHeader h = *reinterpret_cast<const ColHeader *>(a.model);
    float row[N], goal[N];
    const float *gr = a.goal + (p / a.rows_per_goal) * N;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        row[j] = a.rows[p * N + j];
        goal[j] = __ldg(gr + j);
    }
    double th[N];
#pragma unroll
    for (int j = 0; j < N; ++j) th[j] = (double)row[j];
    bool hit = collides<N>(rb, h, a.model, th, box_sm + threadIdx.x);
    int it = 0;
    if (hit) {
        for (; it < a.max_iter;) {
            // float32 throughout, every operation rounded on its own like NumPy's:
            // gradient = gain * ((row - goal) * 1.0);  row = row - 0.01 * gradient
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const float g = __fmul_rn(a.gain, __fsub_rn(row[j], goal[j]));
                row[j] = __fsub_rn(row[j], __fmul_rn(a.step, g));
                th[j] = (double)row[j];
            }
            ++it;
            hit = collides<N>(rb, h, a.model, th, box_sm + threadIdx.x);
            if (!hit) break;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) a.rows[p * N + j] = row[j];
    }
    if (a.iters) a.iters[p] = it;
    if (a.flags) a.flags[p] = hit ? 1 : 0;
}

// URDF.link_fk_batch: (P, L, 4, 4) float64.  A row's L poses are 128 L contiguous bytes, so a thread storing
// its own poses word by word touches 32 different lines per warp instruction (measured: 1.13 TB/s).  Each
// link's pose is staged per warp in shared memory instead (row stride 18 doubles: 16-byte aligned rows,
// conflict-free 16-byte reads) and flushed with 16-byte streaming stores, eight lanes per 128-byte pose:
// every store instruction writes four full lines.
template <int N>
__global__ void __launch_bounds__(kColThreads) link_fk_kernel(const __grid_constant__ RobotPack<double, N> rb,
                                                              const ColArgs a) {
    constexpr int kStride = 18;
    __shared__ __align__(16) double stage_sm[kColThreads / 32][32 * kStride];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t pw = (int64_t)blockIdx.x * kColThreads + warp * 32;
    if (pw >= a.P) return;
    const int64_t rem = a.P - pw;
    const int rows = (int)(rem < 32 ? rem : 32);
    const int64_t p = lane < rows ? pw + lane : pw;  // surplus lanes of the last warp shadow its first row
    const ColHeader h = *reinterpret_cast<const ColHeader *>(a.model);
    const int *lj = reinterpret_cast<const int *>(a.model + h.off_link_joint);
    const double *C = reinterpret_cast<const double *>(a.model + h.off_link_C);
    double th[N];
    load_row<N>(a.theta, a.theta_dtype, false, p, th);
    JointCS<double, N> q;
    joint_cs<double, N, false>(rb, th, q);
    double *o = stage_sm[warp] + lane * kStride;
    Chain<N> ch;
    ch.start(rb);
    // links are visited joint by joint (k = -1 first): a link's pose needs the chain up to its joint only
    for (int k = -1; k < N; ++k) {
        if (k >= 0) {
            // (the joint index is a loop variable here, not an unrolled constant: constant-bank arrays are
            // indexed dynamically, which is fine off the hot path)
            ch.step(rb, k, q.c[k], q.s[k], q.d[k]);
        }
        for (unsigned l = 0; l < h.L; ++l) {
            if (__ldg(lj + l) != k) continue;
            const double *c = C + 12 * l;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double x = k < 0 ? (r == 0) : ch.X[r], y = k < 0 ? (r == 1) : ch.Y[r], z = k < 0 ? (r == 2) : ch.Z[r];
                const double pr = k < 0 ? 0.0 : ch.p[r];
#pragma unroll
                for (int cidx = 0; cidx < 3; ++cidx)
                    o[4 * r + cidx] = x * __ldg(c + cidx) + y * __ldg(c + 3 + cidx) + z * __ldg(c + 6 + cidx);
                o[4 * r + 3] = pr + x * __ldg(c + 9) + y * __ldg(c + 10) + z * __ldg(c + 11);
            }
            o[12] = 0.0; o[13] = 0.0; o[14] = 0.0; o[15] = 1.0;
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int r = it * 4 + (lane >> 3);
                if (r < rows) {
                    const double2 v = *reinterpret_cast<const double2 *>(stage_sm[warp] + r * kStride + (lane & 7) * 2);
                    __stcs(reinterpret_cast<double2 *>(a.T + ((pw + r) * (int64_t)h.L + l) * 16) + (lane & 7), v);
                }
            }
            __syncwarp();
        }
    }
}

static int check_model(const mpk_robot *rb, const void *model_host_header, ColHeader &h) {
    std::memcpy(&h, model_host_header, sizeof h);
    if (h.magic != kColMagic) return fail(MPK_EINVAL, "not a packed collision model");
    if ((int)h.n != rb->n) return fail(MPK_EINVAL, "collision model was packed for another robot");
    return MPK_OK;
}

}  // namespace mpk

using namespace mpk;

// ---- host: pack ---------------------------------------------------------------------------
static void se3_from(const double *T16, double *R, double *p) {
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R[3 * r + c] = T16[4 * r + c];
        p[r] = T16[4 * r + 3];
    }
}

extern "C" size_t mpk_collision_model_bytes(int n, int L, int H, int64_t npoints) {
    if (n < 1 || L < 0 || H < 0 || npoints < 0) return 0;
    size_t b = sizeof(ColHeader);
    auto pad = [](size_t x) { return (x + 7) & ~size_t(7); };
    b += pad(sizeof(int) * L) + sizeof(double) * 12 * L + pad(sizeof(int) * (n + 2)) + pad(sizeof(int) * (H + 1)) +
         pad((size_t)H * H * 2) + sizeof(double) * 3 * npoints;
    return b;
}

extern "C" int mpk_collision_model_pack(const mpk_robot *rb, int L, const int32_t *link_joint, const double *link_home,
                                        const uint8_t *acm, int H, const int32_t *hull_link, const int32_t *hull_count,
                                        const double *hull_points, void *out, size_t out_bytes) {
    if (!rb || !link_joint || !link_home || !out || L < 1 || H < 0) return fail(MPK_EINVAL, "bad collision model arguments");
    if (H > 0 && (!hull_link || !hull_count || !hull_points || !acm))
        return fail(MPK_EINVAL, "hull arrays and the allowed-collision matrix are required");
    if (H > 255) return fail(MPK_EUNSUPPORTED, "at most 255 hulls");
    const int n = rb->n;
    int64_t npoints = 0;
    for (int q = 0; q < H; ++q) {
        if (hull_link[q] < 0 || hull_link[q] >= L || hull_count[q] < 1) return fail(MPK_EINVAL, "bad hull table");
        npoints += hull_count[q];
    }
    for (int l = 0; l < L; ++l)
        if (link_joint[l] < -1 || link_joint[l] >= n) return fail(MPK_EINVAL, "link_joint out of range");
    if (out_bytes < mpk_collision_model_bytes(n, L, H, npoints)) return fail(MPK_EINVAL, "output buffer too small");
    auto pad = [](size_t x) { return (x + 7) & ~size_t(7); };
    unsigned char *m = static_cast<unsigned char *>(out);
    std::memset(m, 0, out_bytes);
    ColHeader h;
    std::memset(&h, 0, sizeof h);
    h.magic = kColMagic;
    h.n = n;
    h.L = L;
    h.H = H;
    h.npoints = (uint32_t)npoints;
    size_t off = sizeof(ColHeader);
    h.off_link_joint = off; off += pad(sizeof(int) * L);
    h.off_link_C = off; off += sizeof(double) * 12 * L;
    h.off_hull_begin = off; off += pad(sizeof(int) * (n + 2));
    h.off_hull_pts = off; off += pad(sizeof(int) * (H + 1));
    h.off_pairs = off; off += pad((size_t)H * H * 2);
    h.off_points = off; off += sizeof(double) * 3 * npoints;
    h.bytes = off;
    int *lj = reinterpret_cast<int *>(m + h.off_link_joint);
    double *C = reinterpret_cast<double *>(m + h.off_link_C);
    // C_l = F_k^-1 T_l(0)   (k = -1: T_l(0) itself)
    for (int l = 0; l < L; ++l) {
        lj[l] = link_joint[l];
        double R[9], p[3];
        se3_from(link_home + 16 * l, R, p);
        double *c = C + 12 * l;
        const int k = link_joint[l];
        if (k < 0) {
            for (int i = 0; i < 9; ++i) c[i] = R[i];
            for (int i = 0; i < 3; ++i) c[9 + i] = p[i];
        } else {
            const double *F = rb->F[k];  // R row-major, p
            for (int r = 0; r < 3; ++r) {
                for (int cc = 0; cc < 3; ++cc) c[3 * r + cc] = F[r] * R[cc] + F[3 + r] * R[3 + cc] + F[6 + r] * R[6 + cc];
                c[9 + r] = F[r] * (p[0] - F[9]) + F[3 + r] * (p[1] - F[10]) + F[6 + r] * (p[2] - F[11]);
            }
        }
    }
    // hulls ordered by joint (stable): the kernel walks the chain once and meets them in that order
    std::vector<int> order;
    int *hb = reinterpret_cast<int *>(m + h.off_hull_begin);
    for (int k = -1; k < n; ++k) {
        hb[k + 1] = (int)order.size();
        for (int q = 0; q < H; ++q)
            if (link_joint[hull_link[q]] == k) order.push_back(q);
    }
    hb[n + 1] = (int)order.size();
    std::vector<int64_t> src_off(H + 1, 0);
    for (int q = 0; q < H; ++q) src_off[q + 1] = src_off[q] + hull_count[q];
    int *hp = reinterpret_cast<int *>(m + h.off_hull_pts);
    double *pts = reinterpret_cast<double *>(m + h.off_points);
    int64_t e = 0;
    for (int s = 0; s < H; ++s) {
        const int q = order[s], l = hull_link[q];
        const double *c = C + 12 * l;
        hp[s] = (int)e;
        for (int v = 0; v < hull_count[q]; ++v, ++e) {
            const double *x = hull_points + 3 * (src_off[q] + v);
            for (int r = 0; r < 3; ++r) pts[3 * e + r] = c[3 * r] * x[0] + c[3 * r + 1] * x[1] + c[3 * r + 2] * x[2] + c[9 + r];
        }
    }
    hp[H] = (int)e;
    // pairs to test: every two hulls whose links are not in the allowed-collision set
    unsigned char *pairs = m + h.off_pairs;
    uint32_t np = 0;
    for (int s = 0; s < H; ++s)
        for (int t = s + 1; t < H; ++t) {
            const int la = hull_link[order[s]], lb = hull_link[order[t]];
            if (la == lb || acm[la * L + lb] || acm[lb * L + la]) continue;
            pairs[2 * np] = (unsigned char)s;
            pairs[2 * np + 1] = (unsigned char)t;
            ++np;
        }
    h.npairs = np;
    std::memcpy(m, &h, sizeof h);
    return MPK_OK;
}

// ---- launchers ----------------------------------------------------------------------------
static int col_common(const mpk_robot *rb, const void *model_header_host, const void *model_dev, int64_t P, ColHeader &h) {
    if (!rb || !model_header_host || !model_dev) return fail(MPK_EINVAL, "robot / model is NULL");
    if (P < 0) return fail(MPK_EINVAL, "negative size");
    return check_model(rb, model_header_host, h);
}

static size_t box_smem(const ColHeader &h) { return sizeof(double) * 6 * (h.H ? h.H : 1) * kColThreads; }

#define MPK_COL_LAUNCH(KERNEL, smem_)                                                                       \
    MPK_DISPATCH_DOF(rb->n, {                                                                               \
        auto kern = KERNEL<N_>;                                                                             \
        if ((smem_) > 32 * 1024)                                                                            \
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_));          \
        kern<<<grid, kColThreads, (smem_), s>>>(narrow<N_>(rb), a);                                         \
    })

extern "C" int mpk_link_fk_batch(const mpk_robot *rb, const void *model_host, const void *model_dev, int64_t P,
                                 const void *theta, int theta_dtype, double *T, void *stream) {
    ColHeader h;
    if (int rc = col_common(rb, model_host, model_dev, P, h)) return rc;
    if (P == 0) return MPK_OK;
    if (!theta || !T) return fail(MPK_EINVAL, "theta and T are required");
    if (!aligned16(T)) return fail(MPK_EINVAL, "T must be 16-byte aligned");
    if (theta_dtype != MPK_F64 && theta_dtype != MPK_F32) return fail(MPK_EINVAL, "bad dtype");
    ColArgs a;
    std::memset(&a, 0, sizeof a);
    a.P = P;
    a.model = static_cast<const unsigned char *>(model_dev);
    a.theta = theta;
    a.theta_dtype = theta_dtype;
    a.T = T;
    const unsigned grid = (unsigned)((P + kColThreads - 1) / kColThreads);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MPK_COL_LAUNCH(link_fk_kernel, (size_t)0);
    return check_launch("link_fk_batch");
}

extern "C" int mpk_self_collision_aabb(const mpk_robot *rb, const void *model_host, const void *model_dev, int64_t P,
                                       const void *theta, int theta_dtype, uint8_t *flags, void *stream) {
    ColHeader h;
    if (int rc = col_common(rb, model_host, model_dev, P, h)) return rc;
    if (P == 0) return MPK_OK;
    if (!theta || !flags) return fail(MPK_EINVAL, "theta and flags are required");
    if (theta_dtype != MPK_F64 && theta_dtype != MPK_F32) return fail(MPK_EINVAL, "bad dtype");
    ColArgs a;
    std::memset(&a, 0, sizeof a);
    a.P = P;
    a.model = static_cast<const unsigned char *>(model_dev);
    a.theta = theta;
    a.theta_dtype = theta_dtype;
    a.flags = flags;
    const unsigned grid = (unsigned)((P + kColThreads - 1) / kColThreads);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t smem = box_smem(h);
    if (smem > 200 * 1024) return fail(MPK_EUNSUPPORTED, "too many hulls for the per-thread box columns");
    MPK_COL_LAUNCH(self_collision_kernel, smem);
    return check_launch("self_collision_aabb");
}

extern "C" int mpk_collision_avoidance(const mpk_robot *rb, const void *model_host, const void *model_dev, int64_t P,
                                       float *rows, const float *goal, int64_t rows_per_goal, double attractive_gain,
                                       double step, int max_iterations, int32_t *iterations, uint8_t *flags,
                                       void *stream) {
    ColHeader h;
    if (int rc = col_common(rb, model_host, model_dev, P, h)) return rc;
    if (P == 0) return MPK_OK;
    if (!rows || !goal || rows_per_goal < 1 || max_iterations < 0)
        return fail(MPK_EINVAL, "rows, goal and a positive rows_per_goal are required");
    ColArgs a;
    std::memset(&a, 0, sizeof a);
    a.P = P;
    a.model = static_cast<const unsigned char *>(model_dev);
    a.rows = rows;
    a.goal = goal;
    a.rows_per_goal = rows_per_goal;
    a.gain = (float)attractive_gain;
    a.step = (float)step;
    a.max_iter = max_iterations;
    a.iters = iterations;
    a.flags = flags;
    const unsigned grid = (unsigned)((P + kColThreads - 1) / kColThreads);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t smem = box_smem(h);
    if (smem > 200 * 1024) return fail(MPK_EUNSUPPORTED, "too many hulls for the per-thread box columns");
    MPK_COL_LAUNCH(collision_avoidance_kernel, smem);
    return check_launch("collision_avoidance");
}
