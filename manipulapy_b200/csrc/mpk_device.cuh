// mpk_device.cuh -- per-thread, register-resident rigid-body algebra for sm_100a.
//
// Every quantity is expressed in JOINT-ALIGNED link frames: frame i is fixed to
// link i, its z axis is the axis of joint i and (for a revolute joint) its
// origin lies on that axis.  In these frames the joint motion is a pure z
// rotation by theta (plus a z translation st*theta for prismatic / helical
// joints), the joint screw is A_i = [0,0,sr, 0,0,st], and the only per-link
// constants are the pose X_i of frame i in frame i-1 at theta_i = 0 and the
// link inertia re-expressed in frame i.  This is mathematically the reference's
// product of exponentials (kinematics/fk.py:61-70, kinematics/jacobian.py:62-73)
// and its link-CoM inertia model (dynamics/mass_matrix.py:66-96) after a
// constant change of frames done once on the host (robot.cu), and costs about
// half the flops of evaluating e^{[S]theta} per joint.
//
// The constant pack is passed to kernels by value as a __grid_constant__
// parameter: with the link loops fully unrolled every constant is a
// constant-bank operand of the FMA that uses it -- no loads, no registers.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MPK_HD __host__ __device__ __forceinline__

namespace mpk {

// Round-to-nearest fp64 / fp32 primitives that the compiler may not contract into FMAs.
// (The host branch exists only so tests/hostcheck can execute the same templates on a CPU.)
#ifdef __CUDA_ARCH__
MPK_HD double rn_mul(double a, double b) { return __dmul_rn(a, b); }
MPK_HD double rn_add(double a, double b) { return __dadd_rn(a, b); }
MPK_HD double rn_sub(double a, double b) { return __dsub_rn(a, b); }
MPK_HD double rn_div(double a, double b) { return __ddiv_rn(a, b); }
MPK_HD float rn_fsub(float a, float b) { return __fsub_rn(a, b); }
#else
MPK_HD double rn_mul(double a, double b) { volatile double r = a * b; return r; }
MPK_HD double rn_add(double a, double b) { volatile double r = a + b; return r; }
MPK_HD double rn_sub(double a, double b) { volatile double r = a - b; return r; }
MPK_HD double rn_div(double a, double b) { volatile double r = a / b; return r; }
MPK_HD float rn_fsub(float a, float b) { volatile float r = a - b; return r; }
#endif

template <typename T, int N>
struct RobotPack {
    T Rx[N][9];  // rotation of frame i in frame i-1 at theta_i = 0 (row-major)
    T px[N][3];  // origin of frame i in frame i-1
    T sr[N];     // 1 revolute / helical, 0 prismatic
    T st[N];     // z translation per unit theta (|v| for a prismatic joint, else 0)
    T I[N][6];   // rigid: rotational inertia about the frame-i origin (xx,xy,xz,yy,yz,zz)
    T h[N][3];   // rigid: mass * centre of mass (in frame i)
    T m[N];      // rigid: mass
    T G[N][21];  // general: upper triangle (row-major) of the symmetric 6x6 inertia in frame i
    T cg[N][3];  // general: origin of the reference's link-CoM frame in frame i
    T mg[N];     // general: G[3,3] in the CoM frame, the mass the reference's gravity term uses
    T Ree[9];    // end-effector home pose in frame n
    T pee[3];
};

// ---- scalar helpers ---------------------------------------------------------
// sin and cos together.  The CUDA library's float64 sincos() is already minimal on its fast
// path (3-FMA reduction + two 7-term polynomials = 22 fp64 instructions, large arguments
// handled by an out-of-line subroutine); a hand-written Cody-Waite + fdlibm-kernel version
// was measured at the same fp64 count with more select instructions and was dropped.
MPK_HD void sincos_t(double x, double *sn, double *cs) {
#ifdef __CUDA_ARCH__
    sincos(x, sn, cs);
#else
    *sn = sin(x);
    *cs = cos(x);
#endif
}
MPK_HD void sincos_t(float x, float *s, float *c) {
#ifdef __CUDA_ARCH__
    sincosf(x, s, c);
#else
    *s = sinf(x);
    *c = cosf(x);
#endif
}

template <typename T, int N>
struct JointCS {
    T c[N], s[N], d[N];
};

// sin/cos (and z offset) of every joint; prismatic joints get the identity rotation.
template <typename T, int N>
MPK_HD void joint_cs(const RobotPack<T, N> &rb, const T (&th)[N],
                                         JointCS<T, N> &q) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (rb.sr[i] != T(0)) {
            sincos_t(th[i], &q.s[i], &q.c[i]);
        } else {
            q.s[i] = T(0);
            q.c[i] = T(1);
        }
        q.d[i] = rb.st[i] * th[i];
    }
}

// Twist (w, v) of frame i-1 coordinates -> frame i coordinates: Ad(T_{i-1,i}^{-1}),
// T_{i-1,i} = X_i * Jz(theta_i).
template <typename T, int N>
MPK_HD void twist_to_child(const RobotPack<T, N> &rb, int i, T c, T s, T d,
                                               T (&w)[3], T (&v)[3]) {
    const T *R = rb.Rx[i];
    const T *p = rb.px[i];
    // u = v + w x p   (written so that every product contracts into an FMA chain)
    const T ux = v[0] + w[1] * p[2] - w[2] * p[1];
    const T uy = v[1] + w[2] * p[0] - w[0] * p[2];
    const T uz = v[2] + w[0] * p[1] - w[1] * p[0];
    // Rx^T *
    const T w1x = R[0] * w[0] + R[3] * w[1] + R[6] * w[2];
    const T w1y = R[1] * w[0] + R[4] * w[1] + R[7] * w[2];
    const T w1z = R[2] * w[0] + R[5] * w[1] + R[8] * w[2];
    T v1x = R[0] * ux + R[3] * uy + R[6] * uz;
    T v1y = R[1] * ux + R[4] * uy + R[7] * uz;
    const T v1z = R[2] * ux + R[5] * uy + R[8] * uz;
    if (rb.st[i] != T(0)) {  // + w1 x (0,0,d)
        v1x += w1y * d;
        v1y -= w1x * d;
    }
    // Rz^T *
    w[0] = c * w1x + s * w1y;
    w[1] = c * w1y - s * w1x;
    w[2] = w1z;
    v[0] = c * v1x + s * v1y;
    v[1] = c * v1y - s * v1x;
    v[2] = v1z;
}

// Rotate a free vector from frame i-1 coordinates to frame i coordinates.
template <typename T, int N>
MPK_HD void vec_to_child(const RobotPack<T, N> &rb, int i, T c, T s, T (&a)[3]) {
    const T *R = rb.Rx[i];
    const T x = R[0] * a[0] + R[3] * a[1] + R[6] * a[2];
    const T y = R[1] * a[0] + R[4] * a[1] + R[7] * a[2];
    const T z = R[2] * a[0] + R[5] * a[1] + R[8] * a[2];
    a[0] = c * x + s * y;
    a[1] = c * y - s * x;
    a[2] = z;
}

// Wrench (n, f) of frame i-1 coordinates -> frame i coordinates (dual of twist_to_child's inverse).
template <typename T, int N>
MPK_HD void wrench_to_child(const RobotPack<T, N> &rb, int i, T c, T s, T d,
                                                T (&n)[3], T (&f)[3]) {
    const T *R = rb.Rx[i];
    const T *p = rb.px[i];
    // u = n - p x f
    const T ux = n[0] - p[1] * f[2] + p[2] * f[1];
    const T uy = n[1] - p[2] * f[0] + p[0] * f[2];
    const T uz = n[2] - p[0] * f[1] + p[1] * f[0];
    const T f1x = R[0] * f[0] + R[3] * f[1] + R[6] * f[2];
    const T f1y = R[1] * f[0] + R[4] * f[1] + R[7] * f[2];
    const T f1z = R[2] * f[0] + R[5] * f[1] + R[8] * f[2];
    T n1x = R[0] * ux + R[3] * uy + R[6] * uz;
    T n1y = R[1] * ux + R[4] * uy + R[7] * uz;
    const T n1z = R[2] * ux + R[5] * uy + R[8] * uz;
    if (rb.st[i] != T(0)) {  // - (0,0,d) x f1
        n1x += d * f1y;
        n1y -= d * f1x;
    }
    n[0] = c * n1x + s * n1y;
    n[1] = c * n1y - s * n1x;
    n[2] = n1z;
    f[0] = c * f1x + s * f1y;
    f[1] = c * f1y - s * f1x;
    f[2] = f1z;
}

// Wrench (n, f) of frame i coordinates -> frame i-1 coordinates, Ad(T_{i-1,i}^{-1})^T, ADDED to
// (an, af):  (an, af) += Ad^T (n, f).  Every term is one link of an FMA chain seeded by an / af.
template <typename T, int N>
MPK_HD void wrench_to_parent_acc(const RobotPack<T, N> &rb, int i, T c, T s, T d, const T (&n)[3],
                                 const T (&f)[3], T (&an)[3], T (&af)[3]) {
    const T *R = rb.Rx[i];
    const T *p = rb.px[i];
    // Rz *
    const T f1x = c * f[0] - s * f[1];
    const T f1y = s * f[0] + c * f[1];
    const T f1z = f[2];
    T n1x = c * n[0] - s * n[1];
    T n1y = s * n[0] + c * n[1];
    const T n1z = n[2];
    if (rb.st[i] != T(0)) {  // + (0,0,d) x f1
        n1x -= d * f1y;
        n1y += d * f1x;
    }
    // Rx *
    const T f2x = R[0] * f1x + R[1] * f1y + R[2] * f1z;
    const T f2y = R[3] * f1x + R[4] * f1y + R[5] * f1z;
    const T f2z = R[6] * f1x + R[7] * f1y + R[8] * f1z;
    an[0] = an[0] + R[0] * n1x + R[1] * n1y + R[2] * n1z + p[1] * f2z - p[2] * f2y;
    an[1] = an[1] + R[3] * n1x + R[4] * n1y + R[5] * n1z + p[2] * f2x - p[0] * f2z;
    an[2] = an[2] + R[6] * n1x + R[7] * n1y + R[8] * n1z + p[0] * f2y - p[1] * f2x;
    af[0] += f2x;
    af[1] += f2y;
    af[2] += f2z;
}

// Wrench (n, f) of frame i coordinates -> frame i-1 coordinates, in place.
template <typename T, int N>
MPK_HD void wrench_to_parent(const RobotPack<T, N> &rb, int i, T c, T s, T d, T (&n)[3], T (&f)[3]) {
    T an[3] = {T(0), T(0), T(0)}, af[3] = {T(0), T(0), T(0)};
    wrench_to_parent_acc(rb, i, c, s, d, n, f, an, af);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        n[k] = an[k];
        f[k] = af[k];
    }
}

// Spatial momentum (n, f) = G_i [w; v].
template <typename T, int N, bool GEN>
MPK_HD void inertia_mul(const RobotPack<T, N> &rb, int i, const T (&w)[3],
                                            const T (&v)[3], T (&n)[3], T (&f)[3]) {
    if (GEN) {
        const T *G = rb.G[i];  // rows: 0:[0..5] 1:[6..10] 2:[11..14] 3:[15..17] 4:[18..19] 5:[20]
        n[0] = G[0] * w[0] + G[1] * w[1] + G[2] * w[2] + G[3] * v[0] + G[4] * v[1] + G[5] * v[2];
        n[1] = G[1] * w[0] + G[6] * w[1] + G[7] * w[2] + G[8] * v[0] + G[9] * v[1] + G[10] * v[2];
        n[2] = G[2] * w[0] + G[7] * w[1] + G[11] * w[2] + G[12] * v[0] + G[13] * v[1] + G[14] * v[2];
        f[0] = G[3] * w[0] + G[8] * w[1] + G[12] * w[2] + G[15] * v[0] + G[16] * v[1] + G[17] * v[2];
        f[1] = G[4] * w[0] + G[9] * w[1] + G[13] * w[2] + G[16] * v[0] + G[18] * v[1] + G[19] * v[2];
        f[2] = G[5] * w[0] + G[10] * w[1] + G[14] * w[2] + G[17] * v[0] + G[19] * v[1] + G[20] * v[2];
    } else {
        const T *I = rb.I[i];
        const T *h = rb.h[i];
        const T m = rb.m[i];
        // n = I w + h x v ;  f = m v + w x h
        n[0] = I[0] * w[0] + I[1] * w[1] + I[2] * w[2] + h[1] * v[2] - h[2] * v[1];
        n[1] = I[1] * w[0] + I[3] * w[1] + I[4] * w[2] + h[2] * v[0] - h[0] * v[2];
        n[2] = I[2] * w[0] + I[4] * w[1] + I[5] * w[2] + h[0] * v[1] - h[1] * v[0];
        f[0] = m * v[0] + w[1] * h[2] - w[2] * h[1];
        f[1] = m * v[1] + w[2] * h[0] - w[0] * h[2];
        f[2] = m * v[2] + w[0] * h[1] - w[1] * h[0];
    }
}

// ---- inverse dynamics -------------------------------------------------------
// Newton-Euler recursion in the joint-aligned frames (SURVEY.md App. C restated
// in those frames).  Equals the reference's  M ddth + c + g + Js^T Ftip
// (dynamics/id_fd.py:38-47) without its finite-difference noise.
//   rigid (GEN = false): gravity enters as a base acceleration [0; -g].
//   general (GEN = true): G_i is any symmetric 6x6; gravity is the reference's explicit
//   wrench [0; G_i[3,3] R_i^T(-g)] at the link-CoM frame origin (dynamics/forces.py:121-131).
//   ftip: space-frame wrench (moment; force) or nullptr.
// Storage of the per-link state the backward pass needs (local wrench of link i, and the
// joint rotation c, s, d of link i+1 that moves link i+1's wrench into frame i):
//   RegStore  : registers (small DOF, and the forward-dynamics path which reuses c, s for CRBA);
//   SmemStore : one shared-memory column per thread (stride = block size, so a warp's
//               accesses are conflict-free); frees 8 (N-1) fp64 registers per thread, which
//               is what lets 20 warps per SM hide the fp64 pipe latency.
// A prismatic joint has c = 1, s = 0 exactly, so SmemStore keeps its z offset d in the s slot.
template <typename T, int N>
struct RegStore {
    T x[N][6];
    JointCS<T, N> q;
    MPK_HD void put(int i, int k, T v) { x[i][k] = v; }
    MPK_HD T get(int i, int k) const { return x[i][k]; }
    MPK_HD void put_cs(const RobotPack<T, N> &, int i, T c, T s, T d) {
        q.c[i] = c;
        q.s[i] = s;
        q.d[i] = d;
    }
    MPK_HD void get_cs(const RobotPack<T, N> &, int i, T &c, T &s, T &d) const {
        c = q.c[i];
        s = q.s[i];
        d = q.d[i];
    }
};
template <typename T, int N, int THREADS>
struct SmemStore {
    T *base;  // shared memory + threadIdx.x; slot l = wrench of link l, (c, s) of link l + 1
    static constexpr int kSlots = N > 1 ? N - 1 : 0;
    static constexpr size_t kBytes = (size_t)kSlots * 8 * THREADS * sizeof(T);
    MPK_HD void put(int i, int k, T v) { base[(i * 8 + k) * THREADS] = v; }
    MPK_HD T get(int i, int k) const { return base[(i * 8 + k) * THREADS]; }
    MPK_HD void put_cs(const RobotPack<T, N> &rb, int i, T c, T s, T d) {
        if (i == 0) return;
        base[((i - 1) * 8 + 6) * THREADS] = c;
        base[((i - 1) * 8 + 7) * THREADS] = rb.sr[i] != T(0) ? s : d;
    }
    MPK_HD void get_cs(const RobotPack<T, N> &rb, int i, T &c, T &s, T &d) const {
        c = base[((i - 1) * 8 + 6) * THREADS];
        const T x = base[((i - 1) * 8 + 7) * THREADS];
        if (rb.sr[i] != T(0)) {
            s = x;
            d = T(0);  // helical joints are rejected by mpk_robot_create
        } else {
            s = T(0);
            d = x;
        }
    }
};

// Joint values straight from register arrays.
template <typename T, int N>
struct ArrayIn {
    const T (&th)[N];
    const T (&dth)[N];
    const T (&ddth)[N];
    MPK_HD void joint(int i, T &a, T &b, T &c) {
        a = th[i];
        b = dth[i];
        c = ddth[i];
    }
};

// `in.joint(i, theta, dtheta, ddtheta)` yields joint i's values when link i is reached, so a
// kernel can produce them lazily (from global memory or from the time scaling) instead of
// holding 3 N values in registers for the whole recursion.
// SYNC: a block barrier at every link boundary keeps the warps of a block within one link of
// each other, so they fetch the same (large, straight-line) code region at the same time and
// share instruction-cache lines instead of each streaming the whole body on its own.
template <typename T, int N, bool GEN, typename In, typename St, bool SYNC = false>
MPK_HD void rnea(const RobotPack<T, N> &rb, In &in, const T (&g)[3], const T *ftip,
                 T (&tau)[N], St &st_) {
    T w[3], v[3], dw[3], dv[3];
    T ag[3];         // general path: -g in the current frame
    T tn[3], tf[3];  // tip wrench carried down to the last frame
    const bool has_tip = ftip != nullptr;
    if (has_tip) {
        tn[0] = ftip[0]; tn[1] = ftip[1]; tn[2] = ftip[2];
        tf[0] = ftip[3]; tf[1] = ftip[4]; tf[2] = ftip[5];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
#ifdef __CUDA_ARCH__
        if (SYNC) __syncthreads();
#endif
        const T sr = rb.sr[i], st = rb.st[i];
        T th_i, qd, qdd, c, s;
        in.joint(i, th_i, qd, qdd);
        if (sr != T(0)) {
            sincos_t(th_i, &s, &c);
        } else {
            s = T(0);
            c = T(1);
        }
        const T d = st * th_i;
        st_.put_cs(rb, i, c, s, d);
        if (i == 0) {
            // base twist is zero and the base acceleration is [0; -g]: V_1 = A_1 qd,
            // dV_1 = [0; R^T(-g)] + A_1 qdd  (ad(V_1) A_1 = 0)
            T a0[3] = {-g[0], -g[1], -g[2]};
            vec_to_child(rb, 0, c, s, a0);
            w[0] = T(0); w[1] = T(0); w[2] = sr * qd;
            v[0] = T(0); v[1] = T(0); v[2] = st * qd;
            dw[0] = T(0); dw[1] = T(0); dw[2] = sr * qdd;
            if (GEN) {
                ag[0] = a0[0]; ag[1] = a0[1]; ag[2] = a0[2];
                dv[0] = T(0); dv[1] = T(0); dv[2] = st * qdd;
            } else {
                dv[0] = a0[0]; dv[1] = a0[1]; dv[2] = a0[2] + st * qdd;
            }
        } else {
            twist_to_child(rb, i, c, s, d, w, v);
            twist_to_child(rb, i, c, s, d, dw, dv);
            if (GEN) vec_to_child(rb, i, c, s, ag);
            // V_i += A_i dth_i ;  dV_i += ad(V_i) A_i dth_i + A_i ddth_i
            w[2] += sr * qd;
            v[2] += st * qd;
            const T a = sr * qd, b = st * qd;
            dw[0] += a * w[1];
            dw[1] -= a * w[0];
            dw[2] += sr * qdd;
            dv[0] = dv[0] + a * v[1] + b * w[1];
            dv[1] = dv[1] - a * v[0] - b * w[0];
            dv[2] += st * qdd;
        }
        if (has_tip) wrench_to_child(rb, i, c, s, d, tn, tf);
        // F_i = G dV - ad(V)^T (G V) = G dV + [w x n + v x f ; w x f]
        T n[3], f[3], dn[3], df[3];
        inertia_mul<T, N, GEN>(rb, i, w, v, n, f);
        inertia_mul<T, N, GEN>(rb, i, dw, dv, dn, df);
        T Fn[3], Ff[3];
        Fn[0] = dn[0] + w[1] * n[2] - w[2] * n[1] + v[1] * f[2] - v[2] * f[1];
        Fn[1] = dn[1] + w[2] * n[0] - w[0] * n[2] + v[2] * f[0] - v[0] * f[2];
        Fn[2] = dn[2] + w[0] * n[1] - w[1] * n[0] + v[0] * f[1] - v[1] * f[0];
        Ff[0] = df[0] + w[1] * f[2] - w[2] * f[1];
        Ff[1] = df[1] + w[2] * f[0] - w[0] * f[2];
        Ff[2] = df[2] + w[0] * f[1] - w[1] * f[0];
        if (GEN) {
            const T mg = rb.mg[i];
            const T *cg = rb.cg[i];
            const T fx = mg * ag[0], fy = mg * ag[1], fz = mg * ag[2];
            Ff[0] += fx;
            Ff[1] += fy;
            Ff[2] += fz;
            Fn[0] = Fn[0] + cg[1] * fz - cg[2] * fy;
            Fn[1] = Fn[1] + cg[2] * fx - cg[0] * fz;
            Fn[2] = Fn[2] + cg[0] * fy - cg[1] * fx;
        }
        if (i < N - 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                st_.put(i, k, Fn[k]);
                st_.put(i, 3 + k, Ff[k]);
            }
        } else {
            // last link: the backward pass starts straight from registers
            T an[3] = {Fn[0], Fn[1], Fn[2]}, af[3] = {Ff[0], Ff[1], Ff[2]};
            if (has_tip) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    an[k] += tn[k];
                    af[k] += tf[k];
                }
            }
            T cj = c, sj = s, dj = d;
#pragma unroll
            for (int j = N - 1; j >= 0; --j) {
                tau[j] = rb.sr[j] * an[2] + rb.st[j] * af[2];
                if (j > 0) {
#ifdef __CUDA_ARCH__
                    if (SYNC) __syncthreads();
#endif
                    // the wrench of link j, moved to frame j-1, is added to link j-1's local wrench
                    if (j < N - 1) st_.get_cs(rb, j, cj, sj, dj);
                    T bn[3] = {st_.get(j - 1, 0), st_.get(j - 1, 1), st_.get(j - 1, 2)};
                    T bf[3] = {st_.get(j - 1, 3), st_.get(j - 1, 4), st_.get(j - 1, 5)};
                    wrench_to_parent_acc(rb, j, cj, sj, dj, an, af, bn, bf);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        an[k] = bn[k];
                        af[k] = bf[k];
                    }
                }
            }
        }
    }
}

// The same recursion with ROLLED link loops (one copy of the link body, `#pragma unroll 1`).
// The fully unrolled form above is ~37 KB of straight-line SASS per pass for N = 6: with a
// dozen unsynchronised warps per SM streaming through it the instruction fetch path (GCC /
// L1.5 instruction cache) saturates -- ncu: gcc instruction requests at 98 % of peak, icc hit
// rate 78 %, `no_instruction` the top stall (profiles/r1_traj_rnea_unrolled.md).  Rolled, the
// whole kernel is a few KB and stays I-cache resident; robot constants are then read from the
// constant bank with a warp-uniform dynamic link index, the backward-pass state comes from
// `st_` and torques leave through `out.put(j, tau_j)` (no dynamically indexed registers).
template <typename T, int N, bool GEN, typename In, typename St, typename Out>
MPK_HD void rnea_rolled(const RobotPack<T, N> &rb, In &in, const T (&g)[3], const T *ftip,
                        St &st_, Out &out) {
    T w[3] = {T(0), T(0), T(0)}, v[3] = {T(0), T(0), T(0)}, dw[3] = {T(0), T(0), T(0)};
    T dv[3], ag[3];
    if (GEN) {
        dv[0] = dv[1] = dv[2] = T(0);
        ag[0] = -g[0]; ag[1] = -g[1]; ag[2] = -g[2];
    } else {
        dv[0] = -g[0]; dv[1] = -g[1]; dv[2] = -g[2];
        ag[0] = ag[1] = ag[2] = T(0);
    }
    T tn[3] = {T(0), T(0), T(0)}, tf[3] = {T(0), T(0), T(0)};
    const bool has_tip = ftip != nullptr;
    if (has_tip) {
        tn[0] = ftip[0]; tn[1] = ftip[1]; tn[2] = ftip[2];
        tf[0] = ftip[3]; tf[1] = ftip[4]; tf[2] = ftip[5];
    }
    T c = T(1), s = T(0), d = T(0);
    T Fn[3], Ff[3];
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        const T sr = rb.sr[i], st = rb.st[i];
        T th_i, qd, qdd;
        in.joint(i, th_i, qd, qdd);
        if (sr != T(0)) {
            sincos_t(th_i, &s, &c);
        } else {
            s = T(0);
            c = T(1);
        }
        d = st * th_i;
        if (i > 0) st_.put_cs(rb, i, c, s, d);
        // (for link 0 the incoming twist is zero; the general transform then costs a few wasted
        // flops but keeps a single copy of the body)
        twist_to_child(rb, i, c, s, d, w, v);
        twist_to_child(rb, i, c, s, d, dw, dv);
        if (GEN) vec_to_child(rb, i, c, s, ag);
        w[2] += sr * qd;
        v[2] += st * qd;
        const T a = sr * qd, b = st * qd;
        dw[0] += a * w[1];
        dw[1] -= a * w[0];
        dw[2] += sr * qdd;
        dv[0] = dv[0] + a * v[1] + b * w[1];
        dv[1] = dv[1] - a * v[0] - b * w[0];
        dv[2] += st * qdd;
        if (has_tip) wrench_to_child(rb, i, c, s, d, tn, tf);
        T n[3], f[3], dn[3], df[3];
        inertia_mul<T, N, GEN>(rb, i, w, v, n, f);
        inertia_mul<T, N, GEN>(rb, i, dw, dv, dn, df);
        Fn[0] = dn[0] + w[1] * n[2] - w[2] * n[1] + v[1] * f[2] - v[2] * f[1];
        Fn[1] = dn[1] + w[2] * n[0] - w[0] * n[2] + v[2] * f[0] - v[0] * f[2];
        Fn[2] = dn[2] + w[0] * n[1] - w[1] * n[0] + v[0] * f[1] - v[1] * f[0];
        Ff[0] = df[0] + w[1] * f[2] - w[2] * f[1];
        Ff[1] = df[1] + w[2] * f[0] - w[0] * f[2];
        Ff[2] = df[2] + w[0] * f[1] - w[1] * f[0];
        if (GEN) {
            const T mg = rb.mg[i];
            const T *cg = rb.cg[i];
            const T fx = mg * ag[0], fy = mg * ag[1], fz = mg * ag[2];
            Ff[0] += fx;
            Ff[1] += fy;
            Ff[2] += fz;
            Fn[0] = Fn[0] + cg[1] * fz - cg[2] * fy;
            Fn[1] = Fn[1] + cg[2] * fx - cg[0] * fz;
            Fn[2] = Fn[2] + cg[0] * fy - cg[1] * fx;
        }
        if (i < N - 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                st_.put(i, k, Fn[k]);
                st_.put(i, 3 + k, Ff[k]);
            }
        }
    }
    // backward pass, starting from the last link's wrench and rotation still in registers
    T an[3] = {Fn[0] + tn[0], Fn[1] + tn[1], Fn[2] + tn[2]};
    T af[3] = {Ff[0] + tf[0], Ff[1] + tf[1], Ff[2] + tf[2]};
#pragma unroll 1
    for (int j = N - 1; j >= 1; --j) {
        out.put(j, rb.sr[j] * an[2] + rb.st[j] * af[2]);
        if (j < N - 1) st_.get_cs(rb, j, c, s, d);
        T bn[3] = {st_.get(j - 1, 0), st_.get(j - 1, 1), st_.get(j - 1, 2)};
        T bf[3] = {st_.get(j - 1, 3), st_.get(j - 1, 4), st_.get(j - 1, 5)};
        wrench_to_parent_acc(rb, j, c, s, d, an, af, bn, bf);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            an[k] = bn[k];
            af[k] = bf[k];
        }
    }
    out.put(0, rb.sr[0] * an[2] + rb.st[0] * af[2]);
}

// Register-resident convenience form over arrays; `q` receives the joint sines / cosines.
template <typename T, int N, bool GEN>
MPK_HD void rnea(const RobotPack<T, N> &rb, const T (&th)[N], const T (&dth)[N], const T (&ddth)[N],
                 const T (&g)[3], const T *ftip, T (&tau)[N], JointCS<T, N> &q) {
    RegStore<T, N> st_;
    ArrayIn<T, N> in{th, dth, ddth};
    rnea<T, N, GEN>(rb, in, g, ftip, tau, st_);
    q = st_.q;
}

// ---- composite rigid body algorithm (rigid inertias) -------------------------
// M[i][j] for j <= i is written to Mm[i][j] AND Mm[j][i].  Matches the reference's
// sym(sum_k J_k^T G_k J_k) (dynamics/mass_matrix.py:62-96).
template <typename T, int N>
MPK_HD void crba(const RobotPack<T, N> &rb, const JointCS<T, N> &q,
                                     T (&Mm)[N][N]) {
    // composite inertia of links i..N-1 in frame i: (I about origin, h = m*com, m)
    T I[6] = {T(0), T(0), T(0), T(0), T(0), T(0)}, h[3] = {T(0), T(0), T(0)}, m = T(0);
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
#pragma unroll
        for (int k = 0; k < 6; ++k) I[k] += rb.I[i][k];
#pragma unroll
        for (int k = 0; k < 3; ++k) h[k] += rb.h[i][k];
        m += rb.m[i];
        // column i: F = Ic A_i
        const T sr = rb.sr[i], st = rb.st[i];
        T n[3], f[3];
        n[0] = sr * I[2] + st * h[1];
        n[1] = sr * I[4] - st * h[0];
        n[2] = sr * I[5];
        f[0] = -sr * h[1];
        f[1] = sr * h[0];
        f[2] = st * m;
        Mm[i][i] = sr * n[2] + st * f[2];
#pragma unroll
        for (int j = i; j > 0; --j) {
            wrench_to_parent(rb, j, q.c[j], q.s[j], q.d[j], n, f);
            const T mij = rb.sr[j - 1] * n[2] + rb.st[j - 1] * f[2];
            Mm[i][j - 1] = mij;
            Mm[j - 1][i] = mij;
        }
        if (i > 0) {
            // re-express the composite in frame i-1: pose (R, p) = X_i * Jz(theta_i)
            const T c = q.c[i], s = q.s[i];
            // rotate by Rz:  I <- Rz I Rz^T, h <- Rz h
            {
                const T cc = c * c, ss = s * s, cs = c * s;
                const T xx = I[0], xy = I[1], xz = I[2], yy = I[3], yz = I[4];
                I[0] = cc * xx - T(2) * cs * xy + ss * yy;
                I[3] = ss * xx + T(2) * cs * xy + cc * yy;
                I[1] = cs * (xx - yy) + (cc - ss) * xy;
                I[2] = c * xz - s * yz;
                I[4] = s * xz + c * yz;
                const T hx = h[0], hy = h[1];
                h[0] = c * hx - s * hy;
                h[1] = s * hx + c * hy;
            }
            if (rb.st[i] != T(0)) {
                // shift origin by p = (0,0,d): I += 2(q.p) 1 - (p q^T + q p^T), q = h + m p / 2
                const T d = q.d[i];
                const T qz = h[2] + T(0.5) * m * d;
                const T qp2 = T(2) * qz * d;
                I[0] += qp2;
                I[3] += qp2;
                I[2] -= d * h[0];
                I[4] -= d * h[1];
                // zz: 2 qz d - 2 d qz = 0
                h[2] += m * d;
            }
            // rotate by Rx: B = R I ; I <- B R^T ; h <- R h
            const T *R = rb.Rx[i];
            const T *p = rb.px[i];
            T Bm[3][3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                Bm[r][0] = R[3 * r] * I[0] + R[3 * r + 1] * I[1] + R[3 * r + 2] * I[2];
                Bm[r][1] = R[3 * r] * I[1] + R[3 * r + 1] * I[3] + R[3 * r + 2] * I[4];
                Bm[r][2] = R[3 * r] * I[2] + R[3 * r + 1] * I[4] + R[3 * r + 2] * I[5];
            }
            T J0 = Bm[0][0] * R[0] + Bm[0][1] * R[1] + Bm[0][2] * R[2];
            T J1 = Bm[0][0] * R[3] + Bm[0][1] * R[4] + Bm[0][2] * R[5];
            T J2 = Bm[0][0] * R[6] + Bm[0][1] * R[7] + Bm[0][2] * R[8];
            T J3 = Bm[1][0] * R[3] + Bm[1][1] * R[4] + Bm[1][2] * R[5];
            T J4 = Bm[1][0] * R[6] + Bm[1][1] * R[7] + Bm[1][2] * R[8];
            T J5 = Bm[2][0] * R[6] + Bm[2][1] * R[7] + Bm[2][2] * R[8];
            const T hx = R[0] * h[0] + R[1] * h[1] + R[2] * h[2];
            const T hy = R[3] * h[0] + R[4] * h[1] + R[5] * h[2];
            const T hz = R[6] * h[0] + R[7] * h[1] + R[8] * h[2];
            // shift origin by p: q = h + m p / 2
            const T hm = T(0.5) * m;
            const T qx = hx + hm * p[0], qy = hy + hm * p[1], qz = hz + hm * p[2];
            const T qp2 = T(2) * (qx * p[0] + qy * p[1] + qz * p[2]);
            I[0] = J0 + qp2 - T(2) * p[0] * qx;
            I[1] = J1 - (p[0] * qy + qx * p[1]);
            I[2] = J2 - (p[0] * qz + qx * p[2]);
            I[3] = J3 + qp2 - T(2) * p[1] * qy;
            I[4] = J4 - (p[1] * qz + qy * p[2]);
            I[5] = J5 + qp2 - T(2) * p[2] * qz;
            h[0] = hx + m * p[0];
            h[1] = hy + m * p[1];
            h[2] = hz + m * p[2];
        }
    }
}

// General inertias: column j of M = rnea(theta, 0, e_j, g = 0), symmetrised like
// dynamics/mass_matrix.py:96.
template <typename T, int N>
MPK_HD void mass_matrix_general(const RobotPack<T, N> &rb, const T (&th)[N], T (&Mm)[N][N]) {
    T zero[N], g0[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int i = 0; i < N; ++i) zero[i] = T(0);
#pragma unroll
    for (int j = 0; j < N; ++j) {
        T e[N], col[N];
#pragma unroll
        for (int i = 0; i < N; ++i) e[i] = (i == j) ? T(1) : T(0);
        JointCS<T, N> q;
        rnea<T, N, true>(rb, th, zero, e, g0, nullptr, col, q);
#pragma unroll
        for (int i = 0; i < N; ++i) Mm[i][j] = col[i];
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i + 1; j < N; ++j) {
            const T a = T(0.5) * (Mm[i][j] + Mm[j][i]);
            Mm[i][j] = a;
            Mm[j][i] = a;
        }
}

template <typename T, int N, bool GEN>
MPK_HD void mass_matrix(const RobotPack<T, N> &rb, const T (&th)[N], const JointCS<T, N> &q,
                        T (&Mm)[N][N]) {
    if (GEN) mass_matrix_general<T, N>(rb, th, Mm);
    else crba<T, N>(rb, q, Mm);
}

// Solve M x = b in place (b <- x) with an unrolled LDL^T; M symmetric positive definite
// (only the lower triangle is read; it is overwritten).
template <typename T, int N>
MPK_HD void ldlt_solve(T (&Mm)[N][N], T (&b)[N]) {
    T dinv[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        T dj = Mm[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) dj -= Mm[j][k] * Mm[j][k] * Mm[k][k];
        Mm[j][j] = dj;
        dinv[j] = T(1) / dj;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            T l = Mm[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) l -= Mm[i][k] * Mm[j][k] * Mm[k][k];
            Mm[i][j] = l * dinv[j];
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int k = 0; k < i; ++k) b[i] -= Mm[i][k] * b[k];
#pragma unroll
    for (int i = 0; i < N; ++i) b[i] *= dinv[i];
#pragma unroll
    for (int i = N - 1; i >= 0; --i)
#pragma unroll
        for (int k = i + 1; k < N; ++k) b[i] -= Mm[k][i] * b[k];
}

// ddtheta = M(theta)^-1 (tau - rnea(theta, dtheta, 0, g, Ftip))  (dynamics/id_fd.py:50-83).
template <typename T, int N, bool GEN>
MPK_HD void forward_dynamics(const RobotPack<T, N> &rb, const T (&th)[N], const T (&dth)[N],
                             const T (&tau)[N], const T (&g)[3], const T *ftip, T (&dd)[N]) {
    JointCS<T, N> q;
    T zero[N], bias[N];
#pragma unroll
    for (int i = 0; i < N; ++i) zero[i] = T(0);
    rnea<T, N, GEN>(rb, th, dth, zero, g, ftip, bias, q);
#pragma unroll
    for (int i = 0; i < N; ++i) dd[i] = tau[i] - bias[i];
    T Mm[N][N];
    mass_matrix<T, N, GEN>(rb, th, q, Mm);
    ldlt_solve<T, N>(Mm, dd);
}

// ---- kinematics ---------------------------------------------------------------
// World pose of frame i accumulated along the chain; Jacobian column i = Ad(T_{0,i}) A_i.
// Tout: row-major 4x4 (16), Jout: row-major (6, N); either may be nullptr.
template <typename T, int N>
MPK_HD void fk_jacobian(const RobotPack<T, N> &rb, const JointCS<T, N> &q,
                                            T *Tout, T *Jout) {
    T R[9] = {T(1), T(0), T(0), T(0), T(1), T(0), T(0), T(0), T(1)};
    T p[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const T *X = rb.Rx[i];
        const T *px = rb.px[i];
        T Rn[9];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            p[r] += R[3 * r] * px[0] + R[3 * r + 1] * px[1] + R[3 * r + 2] * px[2];
#pragma unroll
            for (int cidx = 0; cidx < 3; ++cidx)
                Rn[3 * r + cidx] =
                    R[3 * r] * X[cidx] + R[3 * r + 1] * X[3 + cidx] + R[3 * r + 2] * X[6 + cidx];
        }
        if (Jout) {
            const T sr = rb.sr[i], st = rb.st[i];
            const T zx = Rn[2], zy = Rn[5], zz = Rn[8];
            Jout[0 * N + i] = sr * zx;
            Jout[1 * N + i] = sr * zy;
            Jout[2 * N + i] = sr * zz;
            Jout[3 * N + i] = sr * (p[1] * zz - p[2] * zy) + st * zx;
            Jout[4 * N + i] = sr * (p[2] * zx - p[0] * zz) + st * zy;
            Jout[5 * N + i] = sr * (p[0] * zy - p[1] * zx) + st * zz;
        }
        const T c = q.c[i], s = q.s[i];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            R[3 * r] = c * Rn[3 * r] + s * Rn[3 * r + 1];
            R[3 * r + 1] = c * Rn[3 * r + 1] - s * Rn[3 * r];
            R[3 * r + 2] = Rn[3 * r + 2];
        }
        if (rb.st[i] != T(0)) {
            const T d = q.d[i];
            p[0] += d * R[2];
            p[1] += d * R[5];
            p[2] += d * R[8];
        }
    }
    if (Tout) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int cidx = 0; cidx < 3; ++cidx)
                Tout[4 * r + cidx] = R[3 * r] * rb.Ree[cidx] + R[3 * r + 1] * rb.Ree[3 + cidx] +
                                     R[3 * r + 2] * rb.Ree[6 + cidx];
            Tout[4 * r + 3] =
                p[r] + R[3 * r] * rb.pee[0] + R[3 * r + 1] * rb.pee[1] + R[3 * r + 2] * rb.pee[2];
        }
        Tout[12] = T(0);
        Tout[13] = T(0);
        Tout[14] = T(0);
        Tout[15] = T(1);
    }
}

// ---- time scaling (planning/trajectory.py:15-75) -------------------------------
// float64, the reference's operation order, no FMA contraction (explicit _rn intrinsics),
// so that the single rounding to float32 reproduces the reference bit for bit.
struct TimeScale {
    double s, sd, sdd;
};

MPK_HD TimeScale time_scaling(int64_t idx, int64_t N, double Tf, int method) {
    const double step = rn_div(Tf, (double)(N - 1));
    const double t = rn_mul((double)idx, step);
    const double tau = rn_div(t, Tf);
    TimeScale r;
    if (method == 3) {
        const double tt = rn_mul(tau, tau);
        // s = 3*tau*tau - 2*tau*tau*tau
        r.s = rn_sub(rn_mul(rn_mul(3.0, tau), tau),
                        rn_mul(rn_mul(rn_mul(2.0, tau), tau), tau));
        (void)tt;
        // sd = 6*tau*(1-tau)/Tf
        r.sd = rn_div(rn_mul(rn_mul(6.0, tau), rn_sub(1.0, tau)), Tf);
        // sdd = 6/(Tf*Tf)*(1-2*tau)
        r.sdd = rn_mul(rn_div(6.0, rn_mul(Tf, Tf)), rn_sub(1.0, rn_mul(2.0, tau)));
    } else if (method == 5) {
        const double t2 = rn_mul(tau, tau), t3 = rn_mul(t2, tau), t4 = rn_mul(t2, t2),
                     t5 = rn_mul(t4, tau);
        r.s = rn_add(rn_sub(rn_mul(10.0, t3), rn_mul(15.0, t4)), rn_mul(6.0, t5));
        r.sd = rn_div(
            rn_add(rn_sub(rn_mul(30.0, t2), rn_mul(60.0, t3)), rn_mul(30.0, t4)), Tf);
        r.sdd = rn_div(
            rn_add(rn_sub(rn_mul(60.0, tau), rn_mul(180.0, t2)), rn_mul(120.0, t3)),
            rn_mul(Tf, Tf));
    } else {
        r.s = r.sd = r.sdd = 0.0;
    }
    return r;
}

MPK_HD float clip_f32(float x, float lo, float hi) {
    return x < lo ? lo : (x > hi ? hi : x);
}

// One trajectory sample of joint j: start + s*dth etc, each rounded once to float32.
MPK_HD void traj_point(const TimeScale &ts, double st, double dth, float lo,
                                           float hi, bool clip, float &p, float &v, float &a) {
    p = (float)rn_add(rn_mul(ts.s, dth), st);
    if (clip) p = clip_f32(p, lo, hi);
    v = (float)rn_mul(ts.sd, dth);
    a = (float)rn_mul(ts.sdd, dth);
}

}  // namespace mpk
