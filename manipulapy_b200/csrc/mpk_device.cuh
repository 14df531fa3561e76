// mpk_device.cuh -- per-thread, register-resident rigid-body algebra for sm_100a.
//
// Every quantity is expressed in JOINT-ALIGNED link frames: frame i is fixed to link i, its
// z axis is the axis of joint i and its origin lies on that axis.  The freedom that is left
// (the direction of x_i and the position of the origin along the axis) is spent the way the
// Denavit-Hartenberg convention spends it: x_{i-1} points along the common normal of the
// axes i-1 and i, so that the constant pose of frame i in frame i-1 factors into planar
// rotations and axis-aligned translations,
//
//     T_{i-1,i}(theta_i) = Tx(a_i) Rx(alpha_i) [Ry(beta_i)] Rz(phi_i + theta_i) Tz(d_i)       (revolute)
//                        = Tx(a_i) Rx(alpha_i) [Ry(beta_i)] Rz(phi_i) Tz(d_i + |v| theta_i)   (prismatic)
//
// and moving a twist or a wrench across a joint costs 20-22 fp64 operations instead of the
// 32-35 of a general (R, p) pair.  The optional Ry(beta) factor (Hayati's parametrisation) is
// only non-trivial for consecutive axes that are nearly but not exactly parallel, where the
// common normal is ill-conditioned; robot.cu decides per link, and the kernels skip the
// factor under a warp-uniform test (beta = 0 for every robot shipped with the reference).
// Frame 0 sits in the space frame with a general constant pose (Rb, pb).
//
// This is mathematically the reference's product of exponentials (kinematics/fk.py:61-70,
// kinematics/jacobian.py:62-73) and its link-CoM inertia model (dynamics/mass_matrix.py:66-96)
// after a constant change of frames done once on the host (robot.cu: e^{[S_i] th} F_i =
// F_i Rz(th) for the home pose F_i of frame i), at well under half the flops of evaluating
// e^{[S]theta} per joint.
//
// The constant pack is passed to kernels by value as a __grid_constant__ parameter: with the
// link loops fully unrolled every constant is a constant-bank operand of the FMA that uses
// it -- no loads, no registers.
//
// Template flags used throughout:
//   GEN  general symmetric 6x6 link inertias (else rigid bodies: I about the origin, m c, m)
//   REV  a "plain" chain: every joint revolute (sr = 1, st = 0) and every link plain
//        Denavit-Hartenberg (beta = 0); drops the prismatic and Hayati terms at compile time.
//        Without it joints are told apart by warp-uniform tests on rb.sr / rb.st / rb.sb.
//   The rigid (non-GEN) paths assume a revolute FIRST joint; other chains take the GEN path.
//   HAY  (transform helpers) keep the warp-uniform Ry(beta) step; kernels pass HAY = !REV
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MPK_HD __host__ __device__ __forceinline__
#define MPK_MAX_DOF_ 8  // == MPK_MAX_DOF of include/mpk.h

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
    T Rb[9], pb[3];        // pose of frame 0 in the space frame at theta_0 = 0 (R row-major)
    T a[N];                // i >= 1: length of the common normal, along x_{i-1}
    T ca[N], sa[N];        // i >= 1: twist alpha_i about x_{i-1}
    T cb[N], sb[N];        // i >= 1: Hayati angle beta_i about y (sb = 0: plain Denavit-Hartenberg)
    T phi[N], d[N];        // i >= 1: joint angle offset about z_i and offset along z_i
    T cphi[N], sphi[N];    // cos / sin(phi_i): the constant z rotation of a prismatic joint
    T sr[N];               // 1 revolute, 0 prismatic
    T st[N];               // z translation per unit theta (|v| for a prismatic joint, else 0)
    T I[N][6];             // rigid: rotational inertia about the frame-i origin (xx,xy,xz,yy,yz,zz)
    T h[N][3];             // rigid: mass * centre of mass (in frame i)
    T m[N];                // rigid: mass
    T Ic[N][6];            // rigid: rotational inertia about the centre of mass, frame-i axes
    T com[N][3];           // rigid: centre of mass in frame i
    T G[N][21];            // general: upper triangle (row-major) of the symmetric 6x6 inertia in frame i
    T cg[N][3];            // general: origin of the reference's link-CoM frame in frame i
    T mg[N];               // general: G[3,3] in the CoM frame, the mass the reference's gravity term uses
    T Ree[9], pee[3];      // end-effector home pose in frame n-1
    T trig[17];            // coefficients of sincos_pack (fill_trig_table)
};

// ---- scalar helpers ---------------------------------------------------------
MPK_HD int64_t f64_bits(double x) {
#ifdef __CUDA_ARCH__
    return __double_as_longlong(x);
#else
    int64_t b;
    memcpy(&b, &x, sizeof b);
    return b;
#endif
}
MPK_HD double bits_f64(int64_t b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double(b);
#else
    double x;
    memcpy(&x, &b, sizeof x);
    return x;
#endif
}

MPK_HD double ld_ro(const double *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// Coefficient table of sincos_pack: [0] 2/pi, [1] 1.5 * 2^52, [2..4] -(pi/2) split in three
// (fdlibm's pio2_1, pio2_2, pio2_2t), [5..10] fdlibm __kernel_sin S1..S6, [11..16]
// __kernel_cos C1..C6.
template <typename T>
inline void fill_trig_table(T *t) {
    t[0] = T(6.36619772367581382433e-01);
    t[1] = T(6755399441055744.0);
    t[2] = T(-1.57079632673412561417e+00);
    t[3] = T(-6.07710050630396597660e-11);
    t[4] = T(-2.02226624879595063154e-21);
    t[5] = T(-1.66666666666666324348e-01);
    t[6] = T(8.33333333332248946124e-03);
    t[7] = T(-1.98412698298579493134e-04);
    t[8] = T(2.75573137070700676789e-06);
    t[9] = T(-2.50507602534068634195e-08);
    t[10] = T(1.58969099521155010221e-10);
    t[11] = T(4.16666666666666019037e-02);
    t[12] = T(-1.38888888888741095749e-03);
    t[13] = T(2.48015872894767294178e-05);
    t[14] = T(-2.75573143513906633035e-07);
    t[15] = T(2.08757232129817482790e-09);
    t[16] = T(-1.13596475577881948265e-11);
}

// Library sin and cos together (used for arguments outside sincos_pack's range).
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

// sin and cos of a joint angle, < 1 ulp-ish (fdlibm kernels on [-pi/4, pi/4] after a three-term
// Cody-Waite reduction, exact for |x| < 1e5).  Same 21 fp64 operations as the CUDA library's
// sincos(), but every coefficient is a constant-bank operand of its FMA (the library version
// materialises its 14 coefficients with 28 UMOVs per call) and the quadrant logic is 4 selects
// and 2 sign XORs: ~35 instructions per call instead of ~80, which matters because a 6-joint
// inverse dynamics is only ~1600 instructions.
// (one out-of-line copy per kernel: the library routine is ~400 instructions that the joint
// range never executes)
struct SinCos {
    double s, c;
};
__host__ __device__ __noinline__ inline SinCos sincos_far(double x) {
    SinCos r;
    sincos_t(x, &r.s, &r.c);
    return r;
}

// The reduction + kernels alone, valid for |x| < 1e5 (no branch: several calls in a row form
// one basic block whose independent dependency chains the scheduler interleaves).
template <typename T>
MPK_HD void sincos_near(const T *tc, double x, double *sn, double *cs) {
    const double t = fma(x, (double)tc[0], (double)tc[1]);
    const int k = (int)(uint32_t)f64_bits(t);  // round(x * 2/pi) sits in the low mantissa bits
    const double kd = t - (double)tc[1];
    double r = fma(kd, (double)tc[2], x);
    r = fma(kd, (double)tc[3], r);
    r = fma(kd, (double)tc[4], r);
    const double z = r * r;
    double ps = fma((double)tc[10], z, (double)tc[9]);
    double pc = fma((double)tc[16], z, (double)tc[15]);
    ps = fma(ps, z, (double)tc[8]);
    pc = fma(pc, z, (double)tc[14]);
    ps = fma(ps, z, (double)tc[7]);
    pc = fma(pc, z, (double)tc[13]);
    ps = fma(ps, z, (double)tc[6]);
    pc = fma(pc, z, (double)tc[12]);
    ps = fma(ps, z, (double)tc[5]);
    pc = fma(pc, z, (double)tc[11]);
    const double S = fma(z * r, ps, r);
    const double C = fma(z * z, pc, fma(-0.5, z, 1.0));
    const bool swap = (k & 1) != 0;
    const double s0 = swap ? C : S, c0 = swap ? S : C;
    // quadrant signs: sin flips for k mod 4 in {2, 3}, cos for {1, 2}
    const int64_t ss = (int64_t)(uint64_t)((uint32_t)k & 2u) << 62;
    const int64_t sc = (int64_t)(uint64_t)(((uint32_t)k + 1u) & 2u) << 62;
    *sn = bits_f64(f64_bits(s0) ^ ss);
    *cs = bits_f64(f64_bits(c0) ^ sc);
}
MPK_HD bool sincos_is_near(double x) { return fabs(x) < 1e5; }  // false for NaN / inf too
MPK_HD bool sincos_is_near(float) { return false; }

template <typename T>
MPK_HD void sincos_pack(const T *tc, double x, double *sn, double *cs) {
    if (!sincos_is_near(x)) {
        const SinCos far = sincos_far(x);  // (by value: the outputs never have their address taken)
        *sn = far.s;
        *cs = far.c;
        return;
    }
    sincos_near(tc, x, sn, cs);
}
MPK_HD void sincos_near(const float *, float x, float *s, float *c) { sincos_t(x, s, c); }
MPK_HD void sincos_pack(const float *, float x, float *s, float *c) { sincos_t(x, s, c); }

// Planar rotation of the pair (p, q) by the angle whose cosine / sine are (c, s):
//   rot  : p' = c p - s q,  q' = s p + c q      (Rz on (x, y); Rx on (y, z); Ry on (z, x))
//   rot_t: the inverse rotation
template <typename T>
MPK_HD void rot(T c, T s, T &p, T &q) {
    const T p1 = c * p - s * q;
    q = s * p + c * q;
    p = p1;
}
template <typename T>
MPK_HD void rot_t(T c, T s, T &p, T &q) {
    const T p1 = c * p + s * q;
    q = c * q - s * p;
    p = p1;
}

template <typename T, int N>
struct JointCS {
    T c[N], s[N], d[N];  // cos / sin of (phi_i + theta_i) and the total z offset d_i + st_i theta_i
};

// Joint rotation and z offset of joint i at joint value th.
template <typename T, int N, bool REV>
MPK_HD void joint_rot(const RobotPack<T, N> &rb, int i, T th, T &c, T &s, T &dz) {
    if (REV) {
        sincos_pack(rb.trig, rb.phi[i] + th, &s, &c);
        dz = rb.d[i];
    } else {
        if (rb.sr[i] != T(0)) {
            sincos_pack(rb.trig, rb.phi[i] + th, &s, &c);
        } else {
            c = rb.cphi[i];
            s = rb.sphi[i];
        }
        dz = rb.d[i] + rb.st[i] * th;
    }
}

template <typename T, int N, bool REV = false>
MPK_HD void joint_cs(const RobotPack<T, N> &rb, const T (&th)[N], JointCS<T, N> &q) {
#pragma unroll
    for (int i = 0; i < N; ++i) joint_rot<T, N, REV>(rb, i, th[i], q.c[i], q.s[i], q.d[i]);
}

// All joint rotations up front with ONE range test for the whole chain: the N sin / cos
// evaluations then sit in a single basic block (N independent ~35-instruction dependency chains)
// instead of N blocks each behind its own branch.  Same values as joint_cs().  Used where few
// warps are resident and instruction-level parallelism is what hides the fp64 latency (the
// forward-dynamics rollouts: 2 warps per scheduler).
template <typename T, int N, bool REV>
MPK_HD void joint_cs_all(const RobotPack<T, N> &rb, const T (&th)[N], JointCS<T, N> &q) {
    bool near = true;
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (REV || rb.sr[i] != T(0)) near = near && sincos_is_near(rb.phi[i] + th[i]);
    if (near) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (REV || rb.sr[i] != T(0)) {
                sincos_near(rb.trig, rb.phi[i] + th[i], &q.s[i], &q.c[i]);
                q.d[i] = rb.d[i];
            } else {
                q.c[i] = rb.cphi[i];
                q.s[i] = rb.sphi[i];
                q.d[i] = rb.d[i] + rb.st[i] * th[i];
            }
        }
    } else {
        joint_cs<T, N, REV>(rb, th, q);
    }
}

// -g expressed in frame 0 at theta_0 = 0 (uniform over a batch: launchers compute it once on
// the host and pass it as a kernel argument).
template <typename T, int N>
MPK_HD void base_gravity(const RobotPack<T, N> &rb, const T *g, T (&g0)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) g0[k] = -(rb.Rb[k] * g[0] + rb.Rb[3 + k] * g[1] + rb.Rb[6 + k] * g[2]);
}

// ---- link geometry classes -----------------------------------------------------------------
// Most industrial arms are built from consecutive joint axes that are exactly parallel or
// perpendicular (to within the digits their URDFs carry), that often intersect (a_i = 0) and whose
// link frames often need no offset along the axis (d_i = 0).  robot.cu classifies every link once
// on the host; the signature GEO (4 bits per link i >= 1, at bit 4 i) is a TEMPLATE parameter of the
// plain-chain kernels, so the products with the known 0 / 1 entries of Rx(alpha_i) and the
// a_i / d_i shifts are not compiled at all (csrc/dyn_kernels.cuh instantiates the signatures of
// the arm families found in the reference's robot database; any other robot runs GEO = 0, the
// general code).  Bits of a link's class:
//   kGeoPerp  sin(alpha) == 1 exactly (cos(alpha) is whatever the URDF's truncated pi/2 left, ~1e-10):
//             Rx^T(alpha)(y, z) = (ca y + z, ca z - y): two FMAs instead of two DMUL + two DFMA
//   kGeoPar   alpha == 0 exactly: Rx(alpha) is the identity
//   kGeoA0    a_i == 0 (intersecting or coincident axes)     kGeoD0    d_i == 0
// Skipping a multiplication by an exact 1 or an addition of an exact 0 does not change the result
// of a finite computation, so a specialised kernel returns the general kernel's values (PERP links:
// to the last rounding of the ~1e-10 term).
constexpr unsigned kGeoPerp = 1u, kGeoPar = 2u, kGeoA0 = 4u, kGeoD0 = 8u;
constexpr MPK_HD unsigned geo_class(unsigned geo, int i) { return (geo >> (4 * i)) & 15u; }

// Rx(alpha_i)^T applied to the pair (y, z) of link i's class
template <unsigned GEO, typename T, int N>
MPK_HD void rot_alpha_t(const RobotPack<T, N> &rb, int i, T &y, T &z) {
    const unsigned cls = geo_class(GEO, i);
    if (cls & kGeoPar) return;
    if (cls & kGeoPerp) {
        const T y1 = rb.ca[i] * y + z;
        z = rb.ca[i] * z - y;
        y = y1;
        return;
    }
    rot_t(rb.ca[i], rb.sa[i], y, z);
}

// Rx(alpha_i) applied to the pair (y, z) of link i's class
template <unsigned GEO, typename T, int N>
MPK_HD void rot_alpha(const RobotPack<T, N> &rb, int i, T &y, T &z) {
    const unsigned cls = geo_class(GEO, i);
    if (cls & kGeoPar) return;
    if (cls & kGeoPerp) {
        const T y1 = rb.ca[i] * y - z;
        z = rb.ca[i] * z + y;
        y = y1;
        return;
    }
    rot(rb.ca[i], rb.sa[i], y, z);
}

// ---- moving twists and wrenches across joint i >= 1 -----------------------------
// Twist (w, v) of frame i-1 coordinates -> frame i coordinates, Ad(T_{i-1,i}^{-1}); (c, s, dz)
// from joint_rot.  20 fp64 operations.
template <typename T, int N, bool HAY = true, unsigned GEO = 0>
MPK_HD void twist_to_child(const RobotPack<T, N> &rb, int i, T c, T s, T dz, T (&w)[3], T (&v)[3]) {
    const unsigned cls = geo_class(GEO, i);
    const T a = rb.a[i];
    // Tx(a): the origin moves to a x  =>  v += w x (a, 0, 0)
    T vy = (cls & kGeoA0) ? v[1] : v[1] + a * w[2];
    T vz = (cls & kGeoA0) ? v[2] : v[2] - a * w[1];
    T vx = v[0], wx = w[0], wy = w[1], wz = w[2];
    rot_alpha_t<GEO>(rb, i, wy, wz);
    rot_alpha_t<GEO>(rb, i, vy, vz);
    if (HAY && rb.sb[i] != T(0)) {
        rot_t(rb.cb[i], rb.sb[i], wz, wx);
        rot_t(rb.cb[i], rb.sb[i], vz, vx);
    }
    rot_t(c, s, wx, wy);
    rot_t(c, s, vx, vy);
    // Tz(dz): v += w x (0, 0, dz)
    v[0] = (cls & kGeoD0) ? vx : vx + wy * dz;
    v[1] = (cls & kGeoD0) ? vy : vy - wx * dz;
    v[2] = vz;
    w[0] = wx;
    w[1] = wy;
    w[2] = wz;
}

// The same for the twist of a revolute link 0 at rest in translation: w = (0, 0, wz), v = 0
// (10 operations; products with the known zeros are not something the compiler may drop).
template <typename T, int N, bool HAY = true, unsigned GEO = 0>
MPK_HD void twist_to_child_z(const RobotPack<T, N> &rb, int i, T c, T s, T dz, T wz0, T (&w)[3],
                             T (&v)[3]) {
    if (HAY && rb.sb[i] != T(0)) {
        w[0] = T(0); w[1] = T(0); w[2] = wz0;
        v[0] = T(0); v[1] = T(0); v[2] = T(0);
        twist_to_child(rb, i, c, s, dz, w, v);
        return;
    }
    const unsigned cls = geo_class(GEO, i);
    const T a = rb.a[i];
    if (cls & kGeoPar) {
        // parallel axes: the twist keeps its direction, only the origin moves
        w[0] = T(0); w[1] = T(0); w[2] = wz0;
        const T vy1 = (cls & kGeoA0) ? T(0) : a * wz0;
        v[0] = s * vy1;
        v[1] = c * vy1;
        v[2] = T(0);
        return;
    }
    const T wy1 = (cls & kGeoPerp) ? wz0 : rb.sa[i] * wz0, wz1 = rb.ca[i] * wz0;
    w[0] = s * wy1;
    w[1] = c * wy1;
    w[2] = wz1;
    if (cls & kGeoA0) {
        v[0] = (cls & kGeoD0) ? T(0) : w[1] * dz;
        v[1] = (cls & kGeoD0) ? T(0) : -(w[0] * dz);
        v[2] = T(0);
        return;
    }
    const T vy1 = a * wz1, vz1 = -(a * wy1);
    v[0] = (cls & kGeoD0) ? s * vy1 : s * vy1 + w[1] * dz;
    v[1] = (cls & kGeoD0) ? c * vy1 : c * vy1 - w[0] * dz;
    v[2] = vz1;
}

// ... and for its acceleration: dw = (0, 0, dwz), dv full (15 operations).
template <typename T, int N, bool HAY = true, unsigned GEO = 0>
MPK_HD void accel_to_child_z(const RobotPack<T, N> &rb, int i, T c, T s, T dz, T dwz0,
                             const T (&dv0)[3], T (&dw)[3], T (&dv)[3]) {
    if (HAY && rb.sb[i] != T(0)) {
        dw[0] = T(0); dw[1] = T(0); dw[2] = dwz0;
        dv[0] = dv0[0]; dv[1] = dv0[1]; dv[2] = dv0[2];
        twist_to_child(rb, i, c, s, dz, dw, dv);
        return;
    }
    const unsigned cls = geo_class(GEO, i);
    T vy = (cls & kGeoA0) ? dv0[1] : dv0[1] + rb.a[i] * dwz0, vz = dv0[2];
    rot_alpha_t<GEO>(rb, i, vy, vz);
    if (cls & kGeoPar) {
        dw[0] = T(0); dw[1] = T(0); dw[2] = dwz0;
        dv[0] = c * dv0[0] + s * vy;
        dv[1] = c * vy - s * dv0[0];
        dv[2] = vz;
        return;
    }
    const T wy1 = (cls & kGeoPerp) ? dwz0 : rb.sa[i] * dwz0, wz1 = rb.ca[i] * dwz0;
    dw[0] = s * wy1;
    dw[1] = c * wy1;
    dw[2] = wz1;
    dv[0] = (cls & kGeoD0) ? c * dv0[0] + s * vy : c * dv0[0] + s * vy + dw[1] * dz;
    dv[1] = (cls & kGeoD0) ? c * vy - s * dv0[0] : c * vy - s * dv0[0] - dw[0] * dz;
    dv[2] = vz;
}

// Rotate a free vector from frame i-1 coordinates to frame i coordinates.
template <typename T, int N, bool HAY = true, unsigned GEO = 0>
MPK_HD void vec_to_child(const RobotPack<T, N> &rb, int i, T c, T s, T (&u)[3]) {
    rot_alpha_t<GEO>(rb, i, u[1], u[2]);
    if (HAY && rb.sb[i] != T(0)) rot_t(rb.cb[i], rb.sb[i], u[2], u[0]);
    rot_t(c, s, u[0], u[1]);
}

// Wrench (n, f) of frame i-1 coordinates -> frame i coordinates.
template <typename T, int N, bool HAY = true, unsigned GEO = 0>
MPK_HD void wrench_to_child(const RobotPack<T, N> &rb, int i, T c, T s, T dz, T (&n)[3], T (&f)[3]) {
    const unsigned cls = geo_class(GEO, i);
    const T a = rb.a[i];
    // Tx(a): n -= (a, 0, 0) x f
    if (!(cls & kGeoA0)) {
        n[1] += a * f[2];
        n[2] -= a * f[1];
    }
    rot_alpha_t<GEO>(rb, i, n[1], n[2]);
    rot_alpha_t<GEO>(rb, i, f[1], f[2]);
    if (HAY && rb.sb[i] != T(0)) {
        rot_t(rb.cb[i], rb.sb[i], n[2], n[0]);
        rot_t(rb.cb[i], rb.sb[i], f[2], f[0]);
    }
    rot_t(c, s, n[0], n[1]);
    rot_t(c, s, f[0], f[1]);
    // Tz(dz): n -= (0, 0, dz) x f
    if (!(cls & kGeoD0)) {
        n[0] += dz * f[1];
        n[1] -= dz * f[0];
    }
}

// Space-frame wrench -> frame 0 coordinates (general base pose, then the joint rotation).
template <typename T, int N>
MPK_HD void wrench_to_base(const RobotPack<T, N> &rb, T c, T s, T dz, T (&n)[3], T (&f)[3]) {
    const T *R = rb.Rb;
    const T *p = rb.pb;
    const T ux = n[0] - p[1] * f[2] + p[2] * f[1];
    const T uy = n[1] - p[2] * f[0] + p[0] * f[2];
    const T uz = n[2] - p[0] * f[1] + p[1] * f[0];
    const T fx = R[0] * f[0] + R[3] * f[1] + R[6] * f[2];
    const T fy = R[1] * f[0] + R[4] * f[1] + R[7] * f[2];
    const T fz = R[2] * f[0] + R[5] * f[1] + R[8] * f[2];
    n[0] = R[0] * ux + R[3] * uy + R[6] * uz;
    n[1] = R[1] * ux + R[4] * uy + R[7] * uz;
    n[2] = R[2] * ux + R[5] * uy + R[8] * uz;
    f[0] = fx; f[1] = fy; f[2] = fz;
    rot_t(c, s, n[0], n[1]);
    rot_t(c, s, f[0], f[1]);
    n[0] += dz * f[1];
    n[1] -= dz * f[0];
}

// Wrench (n, f) of frame i coordinates -> frame i-1 coordinates, Ad(T_{i-1,i}^{-1})^T, ADDED to
// (an, af).  22 operations; the rotated terms are links of FMA chains seeded by an / af.
template <typename T, int N, bool HAY = true, unsigned GEO = 0>
MPK_HD void wrench_to_parent_acc(const RobotPack<T, N> &rb, int i, T c, T s, T dz, const T (&n)[3],
                                 const T (&f)[3], T (&an)[3], T (&af)[3]) {
    const unsigned cls = geo_class(GEO, i);
    const T a = rb.a[i], ca = rb.ca[i], sa = rb.sa[i];
    // Tz(dz): n += (0, 0, dz) x f
    T nx = (cls & kGeoD0) ? n[0] : n[0] - dz * f[1];
    T ny = (cls & kGeoD0) ? n[1] : n[1] + dz * f[0];
    T nz = n[2], fx = f[0], fy = f[1], fz = f[2];
    rot(c, s, nx, ny);
    rot(c, s, fx, fy);
    if (HAY && rb.sb[i] != T(0)) {
        rot(rb.cb[i], rb.sb[i], nz, nx);
        rot(rb.cb[i], rb.sb[i], fz, fx);
    }
    // Rx(alpha), then Tx(a): n += (a, 0, 0) x f -- with Tx applied first (they commute)
    if (!(cls & kGeoA0)) {
        ny -= a * fz;
        nz += a * fy;
    }
    an[0] += nx;
    af[0] += fx;
    if (cls & kGeoPar) {
        an[1] += ny;
        an[2] += nz;
        af[1] += fy;
        af[2] += fz;
    } else if (cls & kGeoPerp) {
        an[1] = (an[1] - nz) + ca * ny;
        an[2] = (an[2] + ny) + ca * nz;
        af[1] = (af[1] - fz) + ca * fy;
        af[2] = (af[2] + fy) + ca * fz;
    } else {
        an[1] = an[1] + ca * ny - sa * nz;
        an[2] = an[2] + sa * ny + ca * nz;
        af[1] = af[1] + ca * fy - sa * fz;
        af[2] = af[2] + sa * fy + ca * fz;
    }
}

// Only the z moment of the moved wrench (all a revolute joint i-1 needs): 10 operations.
template <typename T, int N, bool HAY = true, unsigned GEO = 0>
MPK_HD T wrench_to_parent_nz(const RobotPack<T, N> &rb, int i, T c, T s, T dz, const T (&n)[3],
                             const T (&f)[3], T acc) {
    if (HAY && rb.sb[i] != T(0)) {
        T an[3] = {T(0), T(0), acc}, af[3] = {T(0), T(0), T(0)};
        wrench_to_parent_acc(rb, i, c, s, dz, n, f, an, af);
        return an[2];
    }
    const unsigned cls = geo_class(GEO, i);
    const T nx = (cls & kGeoD0) ? n[0] : n[0] - dz * f[1];
    const T ny = (cls & kGeoD0) ? n[1] : n[1] + dz * f[0];
    if (cls & kGeoPar) {
        // only z survives Rx(0): n_z + a f_y'
        if (cls & kGeoA0) return acc + n[2];
        const T fy1 = s * f[0] + c * f[1];
        return acc + (n[2] + rb.a[i] * fy1);
    }
    const T ny1 = (cls & kGeoA0) ? s * nx + c * ny : s * nx + c * ny - rb.a[i] * f[2];
    T nz1 = n[2];
    if (!(cls & kGeoA0)) {
        const T fy1 = s * f[0] + c * f[1];
        nz1 = n[2] + rb.a[i] * fy1;
    }
    if (cls & kGeoPerp) return (acc + ny1) + rb.ca[i] * nz1;
    return acc + rb.sa[i] * ny1 + rb.ca[i] * nz1;
}

// Wrench (n, f) of frame i coordinates -> frame i-1 coordinates, in place.
template <typename T, int N, bool HAY = true, unsigned GEO = 0>
MPK_HD void wrench_to_parent(const RobotPack<T, N> &rb, int i, T c, T s, T dz, T (&n)[3], T (&f)[3]) {
    const unsigned cls = geo_class(GEO, i);
    const T a = rb.a[i];
    if (!(cls & kGeoD0)) {
        n[0] -= dz * f[1];
        n[1] += dz * f[0];
    }
    rot(c, s, n[0], n[1]);
    rot(c, s, f[0], f[1]);
    if (HAY && rb.sb[i] != T(0)) {
        rot(rb.cb[i], rb.sb[i], n[2], n[0]);
        rot(rb.cb[i], rb.sb[i], f[2], f[0]);
    }
    if (!(cls & kGeoA0)) {
        n[1] -= a * f[2];
        n[2] += a * f[1];
    }
    rot_alpha<GEO>(rb, i, n[1], n[2]);
    rot_alpha<GEO>(rb, i, f[1], f[2]);
}

// Spatial momentum (n, f) = G_i [w; v].
template <typename T, int N, bool GEN>
MPK_HD void inertia_mul(const RobotPack<T, N> &rb, int i, const T (&w)[3], const T (&v)[3],
                        T (&n)[3], T (&f)[3]) {
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

// Net wrench (about the frame origin) that moves rigid link i with twist (w, v) and spatial
// acceleration (dw, dv), through the centre of mass c:
//   vc = v + w x c,  ac = dv + dw x c + w x vc,  f = m ac,  n = Ic dw + w x (Ic w) + c x f.
// 51 operations (the spatial-inertia form G dV - ad(V)^T G V needs 66), and only 12 of them
// are FMAs with three register operands -- the ones that run at 2/3 rate on the fp64 pipe.
template <typename T, int N>
MPK_HD void rigid_wrench(const RobotPack<T, N> &rb, int i, const T (&w)[3], const T (&v)[3],
                         const T (&dw)[3], const T (&dv)[3], T (&n)[3], T (&f)[3]) {
    const T *c = rb.com[i];
    const T *J = rb.Ic[i];
    const T m = rb.m[i];
    const T vcx = v[0] + w[1] * c[2] - w[2] * c[1];
    const T vcy = v[1] + w[2] * c[0] - w[0] * c[2];
    const T vcz = v[2] + w[0] * c[1] - w[1] * c[0];
    T ax = dv[0] + dw[1] * c[2] - dw[2] * c[1];
    T ay = dv[1] + dw[2] * c[0] - dw[0] * c[2];
    T az = dv[2] + dw[0] * c[1] - dw[1] * c[0];
    ax = ax + w[1] * vcz - w[2] * vcy;
    ay = ay + w[2] * vcx - w[0] * vcz;
    az = az + w[0] * vcy - w[1] * vcx;
    f[0] = m * ax;
    f[1] = m * ay;
    f[2] = m * az;
    const T lx = J[0] * w[0] + J[1] * w[1] + J[2] * w[2];
    const T ly = J[1] * w[0] + J[3] * w[1] + J[4] * w[2];
    const T lz = J[2] * w[0] + J[4] * w[1] + J[5] * w[2];
    T nx = J[0] * dw[0] + J[1] * dw[1] + J[2] * dw[2];
    T ny = J[1] * dw[0] + J[3] * dw[1] + J[4] * dw[2];
    T nz = J[2] * dw[0] + J[4] * dw[1] + J[5] * dw[2];
    nx = nx + c[1] * f[2] - c[2] * f[1];
    ny = ny + c[2] * f[0] - c[0] * f[2];
    nz = nz + c[0] * f[1] - c[1] * f[0];
    n[0] = nx + w[1] * lz - w[2] * ly;
    n[1] = ny + w[2] * lx - w[0] * lz;
    n[2] = nz + w[0] * ly - w[1] * lx;
}

// ---- inverse dynamics -------------------------------------------------------
// Newton-Euler recursion in the joint-aligned frames (SURVEY.md App. C restated in those
// frames).  Equals the reference's  M ddth + c + g + Js^T Ftip  (dynamics/id_fd.py:38-47)
// without its finite-difference noise.
//   rigid (GEN = false): gravity enters as a base acceleration [0; -g].
//   general (GEN = true): G_i is any symmetric 6x6; gravity is the reference's explicit
//   wrench [0; G_i[3,3] R_i^T(-g)] at the link-CoM frame origin (dynamics/forces.py:121-131).
//   g0  : -g in frame-0 coordinates (base_gravity);  ftip: space-frame wrench (moment; force)
//         or nullptr.
// Storage of the per-link state the backward pass needs (local wrench of link i, and the
// joint rotation of link i+1 that moves link i+1's wrench into frame i):
//   RegStore  : registers (the forward-dynamics path, which reuses c, s for CRBA);
//   SmemStore : one shared-memory column per thread (stride = block size, so a warp's
//               accesses are conflict-free); frees 8 (N-1) fp64 registers per thread, which
//               is what lets 20 warps per SM hide the fp64 pipe latency.
// A prismatic joint's rotation is the constant (cphi, sphi), so SmemStore keeps its variable
// z offset in the s slot instead.
template <typename T, int N>
struct RegStore {
    static constexpr bool kPrecomputedCS = false;
    T x[N][6];
    JointCS<T, N> q;
    MPK_HD void put(int i, int k, T v) { x[i][k] = v; }
    MPK_HD T get(int i, int k) const { return x[i][k]; }
    MPK_HD void put_wrench(int i, const T (&n)[3], const T (&f)[3]) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            x[i][k] = n[k];
            x[i][3 + k] = f[k];
        }
    }
    MPK_HD void get_wrench(int i, T (&n)[3], T (&f)[3]) const {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            n[k] = x[i][k];
            f[k] = x[i][3 + k];
        }
    }
    template <bool REV>
    MPK_HD void put_cs(const RobotPack<T, N> &, int i, T c, T s, T d) {
        q.c[i] = c;
        q.s[i] = s;
        q.d[i] = d;
    }
    template <bool REV>
    MPK_HD void get_cs(const RobotPack<T, N> &, int i, T &c, T &s, T &d) const {
        c = q.c[i];
        s = q.s[i];
        d = q.d[i];
    }
    // (the backward pass asks through its own accessor, so that a store may keep the rotations in
    // registers for the forward pass only)
    template <bool REV>
    MPK_HD void get_cs_back(const RobotPack<T, N> &rb, int i, T &c, T &s, T &d) const {
        get_cs<REV>(rb, i, c, s, d);
    }
};
// ... with the joint rotations q already filled in by the caller (joint_cs_all): rnea() reads
// them instead of evaluating sin / cos link by link.
template <typename T, int N>
struct RegStorePre : RegStore<T, N> {
    static constexpr bool kPrecomputedCS = true;
};
// Whether rnea() takes the shortcut for link 0 (only the z moment about its own axis is
// stored for it): rigid inertias, revolute first joint, at least two links.
// (the rigid kernel flavours are only used for chains whose FIRST joint is revolute)
constexpr bool rnea_fast0(bool GEN, bool REV, int N) { return (void)REV, !GEN && N >= 2; }

// Two values of T side by side: the unit of the link-state column (one 16-byte shared-memory access for doubles).
template <typename T>
struct alignas(2 * sizeof(T)) Pair {
    T a, b;
};
// Volatile pair load: the compiler must not forward the value it stored in the forward pass to the
// backward pass in a register -- that keeps 8 (N-1) doubles alive across the whole recursion,
// which is exactly what this store exists to avoid: 72 -> 122 registers, or spills to local memory.
template <typename T>
MPK_HD Pair<T> ld_pair(const Pair<T> *p) {
#ifdef __CUDA_ARCH__
    Pair<T> r;
    const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
    if constexpr (sizeof(T) == 8) {
        asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.a), "=d"(r.b) : "r"(addr) : "memory");
    } else {
        asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.a), "=f"(r.b) : "r"(addr) : "memory");
    }
    return r;
#else
    return *p;
#endif
}

template <typename T, int N, int THREADS, bool FAST0 = false>
struct SmemStore {
    static constexpr bool kPrecomputedCS = false;
    // shared memory + 2 threadIdx.x (in units of T): a column of PAIRS per thread, pair q of thread t at
    // ((q THREADS + t) pairs: consecutive threads 16 bytes apart, so a warp's access is 512 contiguous bytes.
    // Link l (l >= 1, or every link without FAST0) owns four pairs: (n0, n1) (n2, f0) (f1, f2) of its wrench
    // and (c, s) of link l + 1; with FAST0 link 0 owns two: (its z moment, -) and (c, s) of link 1.
    T *base;
    static constexpr int kPairs = N > 1 ? (FAST0 ? 2 + (N - 2) * 4 : (N - 1) * 4) : 0;
    static constexpr int kValues = 2 * kPairs;
    static constexpr size_t kBytes = (size_t)kValues * THREADS * sizeof(T);
    static constexpr MPK_HD int pair_of(int l, int q) { return FAST0 ? (l == 0 ? q : 2 + (l - 1) * 4 + q) : l * 4 + q; }
    MPK_HD Pair<T> *pp(int l, int q) const { return reinterpret_cast<Pair<T> *>(base) + (size_t)pair_of(l, q) * THREADS; }
    // single values: (i, 2) of link 0 under FAST0 (its z moment) lives in the first half of pair 0
    MPK_HD void put(int i, int k, T v) {
        if (FAST0 && i == 0) pp(0, 0)->a = v;
        else (&pp(i, k >> 1)->a)[k & 1] = v;
    }
    MPK_HD T get(int i, int k) const {
        if (FAST0 && i == 0) return *static_cast<const volatile T *>(&pp(0, 0)->a);
        return *static_cast<const volatile T *>(&pp(i, k >> 1)->a + (k & 1));
    }
    MPK_HD void put_wrench(int i, const T (&n)[3], const T (&f)[3]) {
        *pp(i, 0) = Pair<T>{n[0], n[1]};
        *pp(i, 1) = Pair<T>{n[2], f[0]};
        *pp(i, 2) = Pair<T>{f[1], f[2]};
    }
    MPK_HD void get_wrench(int i, T (&n)[3], T (&f)[3]) const {
        const Pair<T> p0 = ld_pair(pp(i, 0)), p1 = ld_pair(pp(i, 1)), p2 = ld_pair(pp(i, 2));
        n[0] = p0.a; n[1] = p0.b; n[2] = p1.a;
        f[0] = p1.b; f[1] = p2.a; f[2] = p2.b;
    }
    MPK_HD Pair<T> *cs_pair(int i) const { return pp(i - 1, (FAST0 && i == 1) ? 1 : 3); }
    template <bool REV>
    MPK_HD void put_cs(const RobotPack<T, N> &rb, int i, T c, T s, T d) {
        if (i == 0) return;
        *cs_pair(i) = Pair<T>{c, (REV || rb.sr[i] != T(0)) ? s : d};
    }
    template <bool REV>
    MPK_HD void get_cs(const RobotPack<T, N> &rb, int i, T &c, T &s, T &d) const {
        const Pair<T> p = ld_pair(cs_pair(i));
        if (REV || rb.sr[i] != T(0)) {
            c = p.a;
            s = p.b;
            d = rb.d[i];
        } else {
            c = rb.cphi[i];
            s = rb.sphi[i];
            d = p.b;
        }
    }
    template <bool REV>
    MPK_HD void get_cs_back(const RobotPack<T, N> &rb, int i, T &c, T &s, T &d) const {
        get_cs<REV>(rb, i, c, s, d);
    }
};

// ... with the joint rotations evaluated by the caller BEFORE the recursion (all joints behind one
// range test: N independent dependency chains in one basic block, the coefficients fetched once
// instead of once per link).  They go straight to the shared-memory column (where the backward pass
// reads them anyway), so they cost no registers while the recursion runs; only those of the first
// link (consumed at once) and of the last link (needed where the two passes meet) stay in registers.
template <typename T, int N, int THREADS, bool FAST0 = false>
struct SmemStorePre : SmemStore<T, N, THREADS, FAST0> {
    using Base = SmemStore<T, N, THREADS, FAST0>;
    static constexpr bool kPrecomputedCS = true;
    T c0_, s0_, d0_, cl_, sl_, dl_;  // first / last link (d: only touched for chains with a prismatic joint)
    template <bool REV>
    MPK_HD void precompute(const RobotPack<T, N> &rb, const T (&th)[N]) {
        T c[N], s[N];
        bool near = true;
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (REV || rb.sr[i] != T(0)) near = near && sincos_is_near(rb.phi[i] + th[i]);
        if (near) {
#pragma unroll
            for (int i = 0; i < N; ++i)
                if (REV || rb.sr[i] != T(0)) sincos_near(rb.trig, rb.phi[i] + th[i], &s[i], &c[i]);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i)
                if (REV || rb.sr[i] != T(0)) sincos_pack(rb.trig, rb.phi[i] + th[i], &s[i], &c[i]);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            T d = rb.d[i];
            if (!REV && rb.sr[i] == T(0)) {
                c[i] = rb.cphi[i];
                s[i] = rb.sphi[i];
                d = rb.d[i] + rb.st[i] * th[i];
            }
            if (i == 0) {
                c0_ = c[i]; s0_ = s[i]; d0_ = d;
            }
            if (i == N - 1) {
                cl_ = c[i]; sl_ = s[i]; dl_ = d;
            }
            if (i > 0 && i < N - 1) Base::template put_cs<REV>(rb, i, c[i], s[i], d);
        }
    }
    template <bool REV>
    MPK_HD void get_cs(const RobotPack<T, N> &rb, int i, T &c, T &s, T &d) const {
        if (i == 0) {
            c = c0_; s = s0_; d = REV ? rb.d[0] : d0_;
        } else if (i == N - 1) {
            c = cl_; s = sl_; d = REV ? rb.d[N - 1] : dl_;
        } else {
            Base::template get_cs<REV>(rb, i, c, s, d);
        }
    }
};

// Joint values straight from register arrays.
template <typename T, int N>
struct ArrayIn {
    static constexpr bool kZeroAcc = false;
    const T (&th)[N];
    const T (&dth)[N];
    const T (&ddth)[N];
    MPK_HD void joint(int i, T &a, T &b, T &c) {
        a = th[i];
        b = dth[i];
        c = ddth[i];
    }
};
// ... with ddtheta = 0 known at compile time (the bias forces of forward dynamics): rnea() then
// drops the joint-acceleration terms instead of adding zeros.
template <typename T, int N>
struct ArrayInNoAcc {
    static constexpr bool kZeroAcc = true;
    const T (&th)[N];
    const T (&dth)[N];
    MPK_HD void joint(int i, T &a, T &b, T &c) {
        a = th[i];
        b = dth[i];
        c = T(0);
    }
};
// ... with dtheta = ddtheta = 0 known at compile time (gravity forces): every twist is zero and
// the recursion shrinks to rotating the base acceleration down the chain (about 45 operations per
// link instead of 120), which makes it HBM-bound.
template <typename T, int N>
struct ArrayInAtRest {
    static constexpr bool kZeroAcc = true;
    static constexpr bool kZeroVel = true;
    const T (&th)[N];
    MPK_HD void joint(int i, T &a, T &b, T &c) {
        a = th[i];
        b = T(0);
        c = T(0);
    }
};
template <typename In, typename = void>
struct zero_vel_of {
    static constexpr bool value = false;
};
template <typename In>
struct zero_vel_of<In, decltype((void)In::kZeroVel)> {
    static constexpr bool value = In::kZeroVel;
};
template <typename In, typename = void>
struct zero_acc_of {
    static constexpr bool value = false;
};
template <typename In>
struct zero_acc_of<In, decltype((void)In::kZeroAcc)> {
    static constexpr bool value = In::kZeroAcc;
};

// `in.joint(i, theta, dtheta, ddtheta)` yields joint i's values when link i is reached, so a
// kernel can produce them lazily (from global memory or from the time scaling) instead of
// holding 3 N values in registers for the whole recursion.
template <typename T, int N, bool GEN, bool REV, unsigned GEO = 0, typename In, typename St>
MPK_HD void rnea(const RobotPack<T, N> &rb, In &in, const T (&g0)[3], const T *ftip, T (&tau)[N],
                 St &st_) {
    static_assert(GEO == 0 || (REV && !GEN), "link geometry classes exist for plain rigid chains only");
    // rigid inertias and a revolute first joint (guaranteed by the flavour selection): link 0 only
    // contributes the z moment about its own axis, and link 1 receives a twist with known zeros
    constexpr bool FAST0 = rnea_fast0(GEN, REV, N);
    constexpr bool NOACC = zero_acc_of<In>::value;  // ddtheta == 0
    // dtheta == ddtheta == 0 (rigid chains of two or more links; other cases take the full path)
    constexpr bool REST = zero_vel_of<In>::value && FAST0;
    T w[3], v[3], dw[3], dv[3];
    T ag[3];         // general path: -g in the current frame
    T tn[3], tf[3];  // tip wrench carried down to the last frame
    T wz0 = T(0), dwz0 = T(0);
    const bool has_tip = ftip != nullptr;
    if (has_tip) {
        tn[0] = ftip[0]; tn[1] = ftip[1]; tn[2] = ftip[2];
        tf[0] = ftip[3]; tf[1] = ftip[4]; tf[2] = ftip[5];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        T th_i, qd, qdd, c, s, dz;
        in.joint(i, th_i, qd, qdd);
        if constexpr (St::kPrecomputedCS) {
            st_.template get_cs<REV>(rb, i, c, s, dz);
        } else {
            joint_rot<T, N, REV>(rb, i, th_i, c, s, dz);
            st_.template put_cs<REV>(rb, i, c, s, dz);
        }
        if (i == 0) {
            // base twist is zero and the base acceleration is [0; -g]: V_0 = A_0 qd,
            // dV_0 = [0; Rz^T g0] + A_0 qdd  (ad(V_0) A_0 = 0)
            T a0[3] = {g0[0], g0[1], g0[2]};
            rot_t(c, s, a0[0], a0[1]);
            if (has_tip) wrench_to_base(rb, c, s, dz, tn, tf);
            if (FAST0) {
                wz0 = qd;
                dwz0 = qdd;
                dv[0] = a0[0]; dv[1] = a0[1]; dv[2] = a0[2];
                // z moment of link 0's own wrench: (I dw + h x dv)_z (w x I w has no z part
                // for w along z)
                st_.put(0, 2, NOACC ? rb.h[0][0] * a0[1] - rb.h[0][1] * a0[0]
                                    : rb.I[0][5] * qdd + rb.h[0][0] * a0[1] - rb.h[0][1] * a0[0]);
                continue;
            }
            const T sr = REV ? T(1) : rb.sr[0], st = REV ? T(0) : rb.st[0];
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
            if (REST) {
                // no twists: the acceleration of frame i's origin is the rotated base acceleration
                vec_to_child<T, N, !REV, GEO>(rb, i, c, s, dv);
            } else if (FAST0 && i == 1) {
                T dv0[3] = {dv[0], dv[1], dv[2]};
                twist_to_child_z<T, N, !REV, GEO>(rb, 1, c, s, dz, wz0, w, v);
                accel_to_child_z<T, N, !REV, GEO>(rb, 1, c, s, dz, dwz0, dv0, dw, dv);
            } else {
                twist_to_child<T, N, !REV, GEO>(rb, i, c, s, dz, w, v);
                twist_to_child<T, N, !REV, GEO>(rb, i, c, s, dz, dw, dv);
            }
            if (GEN) vec_to_child<T, N, !REV, GEO>(rb, i, c, s, ag);
            if (has_tip) wrench_to_child<T, N, !REV, GEO>(rb, i, c, s, dz, tn, tf);
            // V_i += A_i dth_i ;  dV_i += ad(V_i) A_i dth_i + A_i ddth_i
            if (REST) {
            } else if (REV) {
                w[2] += qd;
                dw[0] += qd * w[1];
                dw[1] -= qd * w[0];
                if (!NOACC) dw[2] += qdd;
                dv[0] += qd * v[1];
                dv[1] -= qd * v[0];
            } else if (rb.sr[i] != T(0)) {  // revolute joint of a mixed chain (warp-uniform test)
                w[2] += qd;
                dw[0] += qd * w[1];
                dw[1] -= qd * w[0];
                if (!NOACC) dw[2] += qdd;
                dv[0] += qd * v[1];
                dv[1] -= qd * v[0];
            } else {  // prismatic: A = [0; 0, 0, st]
                const T b = rb.st[i] * qd;
                v[2] += b;
                dv[0] += b * w[1];
                dv[1] -= b * w[0];
                dv[2] += rb.st[i] * qdd;
            }
        }
        T Fn[3], Ff[3];
        if (GEN) {
            // F_i = G dV - ad(V)^T (G V) = G dV + [w x n + v x f ; w x f]
            T n[3], f[3], dn[3], df[3];
            inertia_mul<T, N, true>(rb, i, w, v, n, f);
            inertia_mul<T, N, true>(rb, i, dw, dv, dn, df);
            Fn[0] = dn[0] + w[1] * n[2] - w[2] * n[1] + v[1] * f[2] - v[2] * f[1];
            Fn[1] = dn[1] + w[2] * n[0] - w[0] * n[2] + v[2] * f[0] - v[0] * f[2];
            Fn[2] = dn[2] + w[0] * n[1] - w[1] * n[0] + v[0] * f[1] - v[1] * f[0];
            Ff[0] = df[0] + w[1] * f[2] - w[2] * f[1];
            Ff[1] = df[1] + w[2] * f[0] - w[0] * f[2];
            Ff[2] = df[2] + w[0] * f[1] - w[1] * f[0];
        } else if (REST) {
            // f = m dv,  n = c x f
            const T *cm = rb.com[i];
            const T m = rb.m[i];
            Ff[0] = m * dv[0];
            Ff[1] = m * dv[1];
            Ff[2] = m * dv[2];
            Fn[0] = cm[1] * Ff[2] - cm[2] * Ff[1];
            Fn[1] = cm[2] * Ff[0] - cm[0] * Ff[2];
            Fn[2] = cm[0] * Ff[1] - cm[1] * Ff[0];
        } else {
            rigid_wrench(rb, i, w, v, dw, dv, Fn, Ff);
        }
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
            st_.put_wrench(i, Fn, Ff);
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
            T cj = c, sj = s, dj = dz;
#pragma unroll
            for (int j = N - 1; j >= 1; --j) {
                tau[j] = (REV || rb.sr[j] != T(0)) ? an[2] : rb.st[j] * af[2];
                // the wrench of link j, moved to frame j-1, is added to link j-1's local wrench
                if (j < N - 1) st_.template get_cs_back<REV>(rb, j, cj, sj, dj);
                if (FAST0 && j == 1) {
                    an[2] = wrench_to_parent_nz<T, N, !REV, GEO>(rb, 1, cj, sj, dj, an, af, st_.get(0, 2));
                } else {
                    T bn[3], bf[3];
                    st_.get_wrench(j - 1, bn, bf);
                    wrench_to_parent_acc<T, N, !REV, GEO>(rb, j, cj, sj, dj, an, af, bn, bf);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        an[k] = bn[k];
                        af[k] = bf[k];
                    }
                }
            }
            tau[0] = (REV || FAST0 || rb.sr[0] != T(0)) ? an[2] : rb.st[0] * af[2];
        }
    }
}

// Register-resident convenience form over arrays; `q` receives the joint rotations.
template <typename T, int N, bool GEN, bool REV, unsigned GEO = 0>
MPK_HD void rnea(const RobotPack<T, N> &rb, const T (&th)[N], const T (&dth)[N], const T (&ddth)[N],
                 const T (&g0)[3], const T *ftip, T (&tau)[N], JointCS<T, N> &q) {
    RegStore<T, N> st_;
    ArrayIn<T, N> in{th, dth, ddth};
    rnea<T, N, GEN, REV, GEO>(rb, in, g0, ftip, tau, st_);
    q = st_.q;
}

// ---- composite rigid body algorithm (rigid inertias) -------------------------
// Rotation of a symmetric 3x3 in the (p, q) plane (r the third axis):  p' = c p - s q,
// q' = s p + c q.
template <typename T>
MPK_HD void rot_inertia(T c, T s, T &pp, T &qq, T &pq, T &pr, T &qr) {
    const T cc = c * c, ss = s * s, cs2 = T(2) * (c * s);
    const T dpq = pp - qq;
    const T pp1 = cc * pp - cs2 * pq + ss * qq;
    const T qq1 = ss * pp + cs2 * pq + cc * qq;
    pq = T(0.5) * cs2 * dpq + (cc - ss) * pq;
    pp = pp1;
    qq = qq1;
    rot(c, s, pr, qr);
}

// Composite inertia (I about the origin, h = m c, m) of frame i coordinates -> frame i-1
// coordinates:  Tz(dz), Rz, [Ry], Rx, Tx(a).  A shift of the coordinates by p maps
// I -> I + 2 (q.p) 1 - (q p^T + p q^T), q = h + m p / 2, and h -> h + m p.
template <typename T, int N, bool HAY = true, unsigned GEO = 0>
MPK_HD void inertia_to_parent(const RobotPack<T, N> &rb, int i, T c, T s, T dz, T (&I)[6], T (&h)[3],
                              T m) {
    const unsigned cls = geo_class(GEO, i);
    if (!(cls & kGeoD0)) {
        const T t = m * dz;
        const T e = (T(2) * h[2] + t) * dz;
        I[0] += e;
        I[3] += e;
        I[2] -= h[0] * dz;
        I[4] -= h[1] * dz;
        h[2] += t;
    }
    rot_inertia(c, s, I[0], I[3], I[1], I[2], I[4]);
    rot(c, s, h[0], h[1]);
    if (HAY && rb.sb[i] != T(0)) {
        rot_inertia(rb.cb[i], rb.sb[i], I[5], I[0], I[2], I[4], I[1]);
        rot(rb.cb[i], rb.sb[i], h[2], h[0]);
    }
    if (!(cls & kGeoPar)) {
        rot_inertia(rb.ca[i], rb.sa[i], I[3], I[5], I[4], I[1], I[2]);
        rot_alpha<GEO>(rb, i, h[1], h[2]);
    }
    if (!(cls & kGeoA0)) {
        const T a = rb.a[i];
        const T t = m * a;
        const T e = (T(2) * h[0] + t) * a;
        I[3] += e;
        I[5] += e;
        I[1] -= h[1] * a;
        I[2] -= h[2] * a;
        h[0] += t;
    }
}

// M[i][j] for j <= i is written to Mm[i][j] AND Mm[j][i].  Matches the reference's
// sym(sum_k J_k^T G_k J_k) (dynamics/mass_matrix.py:62-96).  Rigid inertias, first joint revolute.
template <typename T, int N, bool REV, unsigned GEO = 0>
MPK_HD void crba(const RobotPack<T, N> &rb, const JointCS<T, N> &q, T (&Mm)[N][N]) {
    // composite inertia of links i..N-1 in frame i
    T I[6] = {T(0), T(0), T(0), T(0), T(0), T(0)}, h[3] = {T(0), T(0), T(0)}, m = T(0);
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
#pragma unroll
        for (int k = 0; k < 6; ++k) I[k] += rb.I[i][k];
#pragma unroll
        for (int k = 0; k < 3; ++k) h[k] += rb.h[i][k];
        m += rb.m[i];
        // column i: F = Ic A_i
        T n[3], f[3];
        if (REV) {
            n[0] = I[2]; n[1] = I[4]; n[2] = I[5];
            f[0] = -h[1]; f[1] = h[0]; f[2] = T(0);
            Mm[i][i] = n[2];
        } else if (rb.sr[i] != T(0)) {
            n[0] = I[2]; n[1] = I[4]; n[2] = I[5];
            f[0] = -h[1]; f[1] = h[0]; f[2] = T(0);
            Mm[i][i] = n[2];
        } else {
            const T st = rb.st[i];
            n[0] = st * h[1];
            n[1] = -(st * h[0]);
            n[2] = T(0);
            f[0] = T(0);
            f[1] = T(0);
            f[2] = st * m;
            Mm[i][i] = st * f[2];
        }
#pragma unroll
        for (int j = i; j > 0; --j) {
            T mij;
            if (j == 1) {
                // the first joint of a chain routed here is revolute: only the z moment is needed
                mij = wrench_to_parent_nz<T, N, !REV, GEO>(rb, 1, q.c[1], q.s[1], q.d[1], n, f, T(0));
            } else {
                wrench_to_parent<T, N, !REV, GEO>(rb, j, q.c[j], q.s[j], q.d[j], n, f);
                mij = (REV || rb.sr[j - 1] != T(0)) ? n[2] : rb.st[j - 1] * f[2];
            }
            Mm[i][j - 1] = mij;
            Mm[j - 1][i] = mij;
        }
        if (i > 0) inertia_to_parent<T, N, !REV, GEO>(rb, i, q.c[i], q.s[i], q.d[i], I, h, m);
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
        rnea<T, N, true, false>(rb, th, zero, e, g0, nullptr, col, q);
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

template <typename T, int N, bool GEN, bool REV, unsigned GEO = 0>
MPK_HD void mass_matrix(const RobotPack<T, N> &rb, const T (&th)[N], const JointCS<T, N> &q,
                        T (&Mm)[N][N]) {
    if (GEN) mass_matrix_general<T, N>(rb, th, Mm);
    else crba<T, N, REV, GEO>(rb, q, Mm);
}

// Reciprocal of an LDL^T pivot without the division's special-case branch: hardware seed
// (MUFU.RCP64H, ~2^-20) and two Newton steps, <= 1 ulp for normal, finite arguments (pivots of a
// mass matrix are).  Branch-free, so the N pivots do not cut the factorisation into N basic
// blocks and the seed's latency overlaps the column updates.
MPK_HD double rcp_pivot(double d) {
#ifdef __CUDA_ARCH__
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    e = fma(e, e, e);  // e + e^2: third-order last step
    return fma(x, e, x);
#else
    return 1.0 / d;
#endif
}
MPK_HD float rcp_pivot(float d) { return 1.0f / d; }

// Solve M x = b in place (b <- x) with an unrolled LDL^T; M symmetric positive definite
// (only the lower triangle is read; it is overwritten).
// FASTRCP: pivots inverted by rcp_pivot (forward dynamics) instead of an IEEE division (the
// inverse kinematics, whose iterates are compared step by step with the reference's).
// Factorisation M = L D L^T in place (unit lower L below the diagonal, D on it), 1 / D in dinv ...
template <typename T, int N, bool FASTRCP = false>
MPK_HD void ldlt_factor(T (&Mm)[N][N], T (&dinv)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        T dj = Mm[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) dj -= Mm[j][k] * Mm[j][k] * Mm[k][k];
        Mm[j][j] = dj;
        dinv[j] = FASTRCP ? rcp_pivot(dj) : T(1) / dj;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            T l = Mm[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) l -= Mm[i][k] * Mm[j][k] * Mm[k][k];
            Mm[i][j] = l * dinv[j];
        }
    }
}
// ... and the two triangular solves (b <- M^-1 b).
template <typename T, int N>
MPK_HD void ldlt_apply(const T (&Mm)[N][N], const T (&dinv)[N], T (&b)[N]) {
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
template <typename T, int N, bool FASTRCP = false>
MPK_HD void ldlt_solve(T (&Mm)[N][N], T (&b)[N]) {
    T dinv[N];
    ldlt_factor<T, N, FASTRCP>(Mm, dinv);
    ldlt_apply<T, N>(Mm, dinv, b);
}

// ddtheta = M(theta)^-1 (tau - rnea(theta, dtheta, 0, g, Ftip))  (dynamics/id_fd.py:50-83).
// PHASES > 0: the block's warps meet at a barrier between the phases (joint rotations | bias
// forces | mass matrix | solve), so that they walk through the ~45 KB of straight-line code
// together and share its instruction-cache lines (every thread of the block must call).
MPK_HD void phase_barrier() {
#ifdef __CUDA_ARCH__
    __syncthreads();
#endif
}
template <typename T, int N, bool GEN, bool REV, int PHASES = 0, unsigned GEO = 0>
MPK_HD void forward_dynamics(const RobotPack<T, N> &rb, const T (&th)[N], const T (&dth)[N],
                             const T (&tau)[N], const T (&g0)[3], const T *ftip, T (&dd)[N]) {
    T bias[N];
    RegStorePre<T, N> st_;
    joint_cs_all<T, N, REV>(rb, th, st_.q);
    if (PHASES >= 4) phase_barrier();
    ArrayInNoAcc<T, N> in{th, dth};
    rnea<T, N, GEN, REV, GEO>(rb, in, g0, ftip, bias, st_);
#pragma unroll
    for (int i = 0; i < N; ++i) dd[i] = tau[i] - bias[i];
    if (PHASES >= 2) phase_barrier();
    T Mm[N][N];
    mass_matrix<T, N, GEN, REV, GEO>(rb, th, st_.q, Mm);
    if (PHASES >= 3) phase_barrier();
    ldlt_solve<T, N, true>(Mm, dd);
}

// ---- kinematics ---------------------------------------------------------------
// World pose of frame i accumulated along the chain; Jacobian column i = Ad(T_{0,i}) A_i.
// Tout: row-major 4x4 (16), Jout: row-major (6, N); either may be nullptr.
// body: return the BODY Jacobian Ad(T^-1) J_s instead (kinematics/jacobian.py:74-90), T the
// end-effector pose; for a robot built from S' = Ad(M) B this is the reference's
// J_b[:, i] = Ad(e^{-[B_n] th_n} ... e^{-[B_{i+1}] th_{i+1}}) B_i, and T = M prod e^{[B_i] th_i}.
template <typename T, int N>
MPK_HD void fk_jacobian(const RobotPack<T, N> &rb, const JointCS<T, N> &q, T *Tout, T *Jout,
                        bool body = false) {
    // columns of the world rotation of the current frame, and its origin
    T X[3] = {rb.Rb[0], rb.Rb[3], rb.Rb[6]}, Y[3] = {rb.Rb[1], rb.Rb[4], rb.Rb[7]},
      Z[3] = {rb.Rb[2], rb.Rb[5], rb.Rb[8]};
    T p[3] = {rb.pb[0], rb.pb[1], rb.pb[2]};
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (i > 0) {
            const T a = rb.a[i], ca = rb.ca[i], sa = rb.sa[i];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                p[r] += a * X[r];
                rot_t(ca, sa, Y[r], Z[r]);  // R Rx(alpha): Y' = ca Y + sa Z, Z' = ca Z - sa Y
            }
            if (rb.sb[i] != T(0)) {
#pragma unroll
                for (int r = 0; r < 3; ++r) rot_t(rb.cb[i], rb.sb[i], Z[r], X[r]);  // R Ry(beta)
            }
        }
        if (Jout) {
            const T sr = rb.sr[i], st = rb.st[i];
            Jout[0 * N + i] = sr * Z[0];
            Jout[1 * N + i] = sr * Z[1];
            Jout[2 * N + i] = sr * Z[2];
            Jout[3 * N + i] = sr * (p[1] * Z[2] - p[2] * Z[1]) + st * Z[0];
            Jout[4 * N + i] = sr * (p[2] * Z[0] - p[0] * Z[2]) + st * Z[1];
            Jout[5 * N + i] = sr * (p[0] * Z[1] - p[1] * Z[0]) + st * Z[2];
        }
        const T c = q.c[i], s = q.s[i], dz = q.d[i];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            rot_t(c, s, X[r], Y[r]);  // R Rz(psi): X' = c X + s Y, Y' = c Y - s X
            p[r] += dz * Z[r];
        }
    }
    if (Tout || (body && Jout)) {
        T Rf[9], pf[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int cidx = 0; cidx < 3; ++cidx)
                Rf[3 * r + cidx] = X[r] * rb.Ree[cidx] + Y[r] * rb.Ree[3 + cidx] + Z[r] * rb.Ree[6 + cidx];
            pf[r] = p[r] + X[r] * rb.pee[0] + Y[r] * rb.pee[1] + Z[r] * rb.pee[2];
        }
        if (Tout) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int cidx = 0; cidx < 3; ++cidx) Tout[4 * r + cidx] = Rf[3 * r + cidx];
                Tout[4 * r + 3] = pf[r];
            }
            Tout[12] = T(0);
            Tout[13] = T(0);
            Tout[14] = T(0);
            Tout[15] = T(1);
        }
        if (body && Jout) {
            // column by column: w_b = R^T w_s, v_b = R^T (v_s - p x w_s)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const T wx = Jout[0 * N + i], wy = Jout[1 * N + i], wz = Jout[2 * N + i];
                const T ux = Jout[3 * N + i] - (pf[1] * wz - pf[2] * wy);
                const T uy = Jout[4 * N + i] - (pf[2] * wx - pf[0] * wz);
                const T uz = Jout[5 * N + i] - (pf[0] * wy - pf[1] * wx);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    Jout[k * N + i] = Rf[k] * wx + Rf[3 + k] * wy + Rf[6 + k] * wz;
                    Jout[(3 + k) * N + i] = Rf[k] * ux + Rf[3 + k] * uy + Rf[6 + k] * uz;
                }
            }
        }
    }
}

// ---- Cartesian straight-line trajectory (planning/trajectory.py:504-594, 676-740) ----------
// Rotation vector of E = Rstart^T Rend as utils/so3.py:172-191 (MatrixLog3) computes it: angle
// from atan2(|vee|/2, cos) (:150-169), the generic 0.5 (theta / sin theta) vee with its Taylor
// band near the identity (:114-147), and across (pi - 1e-2, pi] the half-turn form theta * n
// with n from the symmetric part (:33-79).
MPK_HD void so3_log_vec(const double (&E)[9], double (&w)[3]) {
    double c = ((E[0] + E[4] + E[8]) - 1.0) / 2.0;
    c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
    const double vee[3] = {E[7] - E[5], E[2] - E[6], E[3] - E[1]};
    const double vv = vee[0] * vee[0] + vee[1] * vee[1] + vee[2] * vee[2];
    const double sin_t = sqrt(vv > 1e-300 ? vv : 1e-300) / 2.0;
    const double theta = atan2(sin_t, c);
    if (theta > 3.141592653589793 - 1e-2) {
        // columns of (E + E^T)/2 - cos(theta) 1 = (1 - cos theta) n n^T are parallel to the axis
        const double s00 = E[0] - c, s11 = E[4] - c, s22 = E[8] - c;
        const double s01 = 0.5 * (E[1] + E[3]), s02 = 0.5 * (E[2] + E[6]), s12 = 0.5 * (E[5] + E[7]);
        double cx, cy, cz, ref;
        if (s22 >= 1e-6) {
            cx = s02; cy = s12; cz = s22; ref = vee[2];
        } else if (s11 >= 1e-6) {
            cx = s01; cy = s11; cz = s12; ref = vee[1];
        } else {
            cx = s00; cy = s01; cz = s02; ref = vee[0];
        }
        const double n2 = cx * cx + cy * cy + cz * cz;
        const double k = (ref >= 0.0 ? theta : -theta) / sqrt(n2 > 1e-24 ? n2 : 1e-24);
        w[0] = k * cx;
        w[1] = k * cy;
        w[2] = k * cz;
        return;
    }
    double coef;
    if (c > 1.0 - 5e-5) {
        const double u = 1.0 - c;
        coef = 1.0 + u / 3.0 + u * u * (4.0 / 45.0);
    } else {
        const double cs = c < -1.0 + 1e-7 ? -1.0 + 1e-7 : (c > 1.0 - 1e-7 ? 1.0 - 1e-7 : c);
        const double d = 1.0 - cs * cs;
        coef = acos(cs) / sqrt(d > 1e-30 ? d : 1e-30);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k] = 0.5 * coef * vee[k];
}

// Step `idx` of the straight-line motion from Xs to Xe (row-major 4x4):
//   R = Rs exp([w] s), p = s pe + (1 - s) ps (trajectory.py:541-556), v = ds (pe - ps),
//   a = dds (pe - ps) (:700-721); s is cubic for method 3 and QUINTIC for anything else (:544-547),
//   ds / dds vanish for a method other than 3 / 5 (:716-717).  float64, one rounding to float32.
MPK_HD void cartesian_point(const double *Xs, const double *Xe, int64_t idx, int64_t N, double Tf, int method,
                            float (&pos)[3], float (&vel)[3], float (&acc)[3], float (&R)[9]) {
    double Rs[9], E[9], w[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Rs[3 * r + c] = ld_ro(Xs + 4 * r + c);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            E[3 * i + j] = Rs[i] * ld_ro(Xe + j) + Rs[3 + i] * ld_ro(Xe + 4 + j) + Rs[6 + i] * ld_ro(Xe + 8 + j);
    so3_log_vec(E, w);
    // time scaling of the orientation / position: utils/time_scaling.py:28-53 at t = timegap * i
    const double tq = rn_div(rn_mul(rn_div(Tf, (double)N - 1.0), (double)idx), Tf);
    double s;
    if (method == 3) s = 3.0 * (tq * tq) - 2.0 * (tq * tq * tq);
    else s = 10.0 * (tq * tq * tq) - 15.0 * (tq * tq * tq * tq) + 6.0 * (tq * tq * tq * tq * tq);
    // ... and of the linear velocity / acceleration (trajectory.py:703-717): idx * (Tf / (N - 1)) / Tf there, the same
    // double as tq (a product does not depend on the order of its factors, and (double)N - 1 == (double)(N - 1))
    const double tau = tq;
    double sd = 0.0, sdd = 0.0;
    if (method == 3) {
        sd = 6.0 * tau * (1.0 - tau) / Tf;
        sdd = 6.0 / (Tf * Tf) * (1.0 - 2.0 * tau);
    } else if (method == 5) {
        const double t2 = tau * tau, t3 = t2 * tau, t4 = t2 * t2;
        sd = (30.0 * t2 - 60.0 * t3 + 30.0 * t4) / Tf;
        sdd = (60.0 * tau - 180.0 * t2 + 120.0 * t3) / (Tf * Tf);
    }
    // Rodrigues with the Taylor-safe coefficients of utils/so3.py:199-219
    const double kx = w[0] * s, ky = w[1] * s, kz = w[2] * s;
    const double th2 = kx * kx + ky * ky + kz * kz;
    double A, B;
    if (th2 < 1e-4) {
        A = 1.0 - th2 / 6.0 + th2 * th2 / 120.0;
        B = 0.5 - th2 / 24.0 + th2 * th2 / 720.0;
    } else {
        const double th = sqrt(th2 > 1e-12 ? th2 : 1e-12);
        double sn, cs;
        sincos_t(th, &sn, &cs);  // (one argument reduction for both)
        A = sn / th;
        B = (1.0 - cs) / (th * th);
    }
    // exp = 1 + A K + B K^2,  K = [k]x,  K^2 = k k^T - |k|^2 1
    const double X[9] = {1.0 + B * (kx * kx - th2), -A * kz + B * kx * ky, A * ky + B * kx * kz,
                         A * kz + B * kx * ky, 1.0 + B * (ky * ky - th2), -A * kx + B * ky * kz,
                         -A * ky + B * kx * kz, A * kx + B * ky * kz, 1.0 + B * (kz * kz - th2)};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            R[3 * r + c] = (float)(Rs[3 * r] * X[c] + Rs[3 * r + 1] * X[3 + c] + Rs[3 * r + 2] * X[6 + c]);
        const double ps = ld_ro(Xs + 4 * r + 3), pe = ld_ro(Xe + 4 * r + 3);
        pos[r] = (float)(s * pe + (1.0 - s) * ps);
        vel[r] = (float)(sd * (pe - ps));
        acc[r] = (float)(sdd * (pe - ps));
    }
}

// ---- damped-least-squares inverse kinematics (kinematics/ik.py:39-311) -------------------
template <typename T, int NMAX>
struct IkParams {
    T eomg, ev, mu, step_cap, w_rot, w_pos;
    T damping;          // lambda (mu = lambda^2 + 1e-12 unless adaptive_tuning moves lambda)
    int max_iter;
    int adaptive;       // adaptive_tuning: Levenberg-Marquardt damping / step-cap adaptation (ik.py:215-229)
    int backtracking;   // five-scale line search (ik.py:253-276)
    T lo[NMAX], hi[NMAX];
};

template <int NMAX = 8>
inline IkParams<double, NMAX> make_ik_params(int n, double eomg, double ev, int max_iterations, double damping,
                                             double step_cap, double w_rot, double w_pos,
                                             const double *joint_limits, int flags = 0) {
    IkParams<double, NMAX> p;
    p.eomg = eomg;
    p.ev = ev;
    p.mu = damping * damping + 1e-12;  // sigma / (sigma^2 + lambda^2 + 1e-12), ik.py:151
    p.damping = damping;
    p.adaptive = flags & 1;
    p.backtracking = (flags >> 1) & 1;
    p.step_cap = step_cap;
    p.w_rot = w_rot;
    p.w_pos = w_pos;
    p.max_iter = max_iterations;
    for (int j = 0; j < NMAX; ++j) {
        p.lo[j] = (joint_limits && j < n) ? joint_limits[2 * j] : -INFINITY;
        p.hi[j] = (joint_limits && j < n) ? joint_limits[2 * j + 1] : INFINITY;
    }
    return p;
}

// Standard normal from a counter: splitmix64 of (seed, target, draw) -> Box-Muller.
MPK_HD double ik_normal(unsigned long long seed, unsigned long long a, unsigned long long b) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (a + 1) + 0xBF58476D1CE4E5B9ULL * (b + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    const double u1 = ((double)(z >> 11) + 1.0) * (1.0 / 9007199254740993.0);
    unsigned long long y = z * 0xD6E8FEB86659FD93ULL + 0x2545F4914F6CDD1DULL;
    y = (y ^ (y >> 32)) * 0xD6E8FEB86659FD93ULL;
    y ^= y >> 32;
    const double u2 = (double)(y >> 11) * (1.0 / 9007199254740992.0);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

// Geometric pose error of ik.py:88-140: V = [R_c w; p_d - p_c], rot = |angle|, trans = |p_d - p_c|;
// Tc: current pose (row-major 4x4), Td: target pose.
MPK_HD void ik_error(const double (&Tc)[16], const double *Td, double (&V)[6], double &rot, double &trans) {
    double Rd[9], E[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Rd[3 * r + c] = ld_ro(Td + 4 * r + c);
        V[3 + r] = ld_ro(Td + 4 * r + 3) - Tc[4 * r + 3];
    }
    trans = sqrt(V[3] * V[3] + V[4] * V[4] + V[5] * V[5]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            E[3 * i + j] = Tc[i] * Rd[j] + Tc[4 + i] * Rd[3 + j] + Tc[8 + i] * Rd[6 + j];  // R_c^T R_d
    double tr = ((E[0] + E[4] + E[8]) - 1.0) / 2.0;
    tr = tr < -1.0 ? -1.0 : (tr > 1.0 ? 1.0 : tr);
    const double angle = acos(tr);
    rot = fabs(angle);
    double w[3];
    const double vee[3] = {E[7] - E[5], E[2] - E[6], E[3] - E[1]};
    if (angle < 1e-6) {
#pragma unroll
        for (int k = 0; k < 3; ++k) w[k] = vee[k] / 2.0;
    } else if (fabs(angle - 3.141592653589793) < 1e-6) {
        int idx = 0;  // argmax of the diagonal (first maximum)
        if (E[4] > E[0]) idx = 1;
        if (E[8] > E[4 * idx]) idx = 2;
#pragma unroll
        for (int k = 0; k < 3; ++k) w[k] = k == idx ? angle : 0.0;
    } else {
        const double den = 2.0 * sin(angle) + 1e-10;
#pragma unroll
        for (int k = 0; k < 3; ++k) w[k] = angle * (vee[k] / den);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) V[r] = Tc[4 * r] * w[0] + Tc[4 * r + 1] * w[1] + Tc[4 * r + 2] * w[2];
}

// Solver state of one target between iteration windows.
template <typename T, int N>
struct IkState {
    T th[N], best[N];
    double best_err;
    int stall, k;
    int restarts;  // stagnation restarts taken so far (row of the caller's noise table to use next)
    // adaptive_tuning (unused otherwise): current lambda and step cap, last error, growth factor
    double damping, step_cap, prev_err, nu;
};

template <typename T, int N>
MPK_HD void ik_state_init(IkState<T, N> &st, const T (&th0)[N], const IkParams<T, MPK_MAX_DOF_> &prm) {
#pragma unroll
    for (int j = 0; j < N; ++j) st.best[j] = st.th[j] = th0[j];
    st.best_err = INFINITY;
    st.stall = 0;
    st.k = 0;
    st.restarts = 0;
    st.damping = prm.damping;
    st.step_cap = prm.step_cap;
    st.prev_err = INFINITY;
    st.nu = 2.0;
}

// rot + trans of the pose reached at joint values th (the line search's trial points)
template <typename T, int N>
MPK_HD double ik_pose_error(const RobotPack<T, N> &rb, const T (&th)[N], const double *Td) {
    JointCS<T, N> q;
    joint_cs(rb, th, q);
    double Tc[16], V[6], rot, trans;
    fk_jacobian<T, N>(rb, q, Tc, (T *)nullptr);
    ik_error(Tc, Td, V, rot, trans);
    return rot + trans;
}

// Iterations st.k .. k_stop-1 of one target (k_stop <= max_iter).  J: scratch for the 6 x N
// Jacobian of the current iterate (the kernel passes the thread's shared-memory row).
// Returns true when the target is FINISHED -- converged, or max_iter reached (then the
// best-iterate fall-back of ik.py:264-275 has been applied): st.th is the answer, `ok` the
// success flag, `iterations` the reference's count (k + 1; max_iter + 1 when exhausted).
// Returns false when k_stop was reached first: st carries everything the next window needs.
// noise: this target's table of standard normals for the stagnation restarts, `noise_rows` rows of
// N (restart r uses row r), or nullptr / exhausted: the counter-based generator keyed by (seed,
// target, iteration).  A caller that fills the table from NumPy's global generator reproduces the
// reference's restarts draw for draw (Python mirror: single-target calls).
template <typename T, int N>
MPK_HD bool ik_dls_window(const RobotPack<T, N> &rb, const double *Td, IkState<T, N> &st,
                          const IkParams<T, MPK_MAX_DOF_> &prm, unsigned long long seed,
                          unsigned long long target, T *J, int k_stop, bool &ok, int &iterations,
                          const double *noise = nullptr, int noise_rows = 0) {
    T (&th)[N] = st.th;
    double cur = INFINITY, rot = 0.0, trans = 0.0;
    int k = st.k;
    ok = false;
    for (; k < k_stop; ++k) {
        JointCS<T, N> q;
        joint_cs(rb, th, q);
        double Tc[16], V[6];
        fk_jacobian<T, N>(rb, q, Tc, J);
        ik_error(Tc, Td, V, rot, trans);
        cur = rot + trans;
        if (rot < prm.eomg && trans < prm.ev) {
            ok = true;
            break;
        }
        if (cur < st.best_err) {
            st.best_err = cur;
#pragma unroll
            for (int j = 0; j < N; ++j) st.best[j] = th[j];
            st.stall = 0;
        } else {
            ++st.stall;
        }
        if (st.stall > 20) {
            // stagnation restart around the best iterate (ik.py:206-213)
            const bool tabled = noise != nullptr && st.restarts < noise_rows;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double z = tabled ? ld_ro(noise + (int64_t)st.restarts * N + j)
                                        : ik_normal(seed, target, (unsigned long long)k * N + j);
                const T x = st.best[j] + 0.1 * z;
                th[j] = fmin(fmax(x, prm.lo[j]), prm.hi[j]);
            }
            ++st.restarts;
            st.stall = 0;
            st.damping = prm.damping;
            st.nu = 2.0;
            continue;
        }
        T mu = prm.mu, cap = prm.step_cap;
        if (prm.adaptive) {
            // Levenberg-Marquardt adaptation of lambda and of the step cap (ik.py:215-229)
            if (k > 0) {
                if (cur < st.prev_err * 0.75) {
                    st.damping = fmax(1e-6, st.damping / 3);
                    st.step_cap = fmin(prm.step_cap * 1.5, st.step_cap * 1.2);
                    st.nu = 2.0;
                } else if (cur < st.prev_err * 0.95) {
                    st.damping = fmax(1e-6, st.damping / 1.5);
                } else if (cur > st.prev_err) {
                    st.damping = fmin(5e-1, st.damping * st.nu);
                    st.nu = fmin(st.nu * 1.5, 8.0);
                    st.step_cap = fmax(0.01, st.step_cap * 0.7);
                }
            }
            st.prev_err = cur;
            mu = (T)(st.damping * st.damping + 1e-12);
            cap = (T)st.step_cap;
        }
        // (J J^T + mu 1) y = W e ;  dtheta = J^T y
        T A[6][6], y[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            y[r] = V[r] * (r < 3 ? prm.w_rot : prm.w_pos);
#pragma unroll
            for (int c = 0; c <= r; ++c) {
                T s = r == c ? mu : T(0);
#pragma unroll
                for (int i = 0; i < N; ++i) s += J[r * N + i] * J[c * N + i];
                A[r][c] = s;
            }
        }
        ldlt_solve<T, 6>(A, y);
        T d[N], nrm = T(0);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            T s = T(0);
#pragma unroll
            for (int r = 0; r < 6; ++r) s += J[r * N + i] * y[r];
            d[i] = s;
            nrm += s * s;
        }
        nrm = sqrt(nrm);
        const T scale = nrm > cap ? cap / nrm : T(1);
        if (prm.backtracking) {
            // line search over five scales of the (capped) step (ik.py:253-276): the best trial
            // point replaces theta when it is not worse than 1.1 x the current error
#pragma unroll
            for (int j = 0; j < N; ++j) d[j] *= scale;
            T bst[N];
            double bst_err = cur;
#pragma unroll
            for (int j = 0; j < N; ++j) bst[j] = th[j];
            for (int t = 0; t < 5; ++t) {
                const T sc = t == 0 ? T(1) : t == 1 ? T(0.5) : t == 2 ? T(0.25) : t == 3 ? T(0.125) : T(0.75);
                T cand[N];
#pragma unroll
                for (int j = 0; j < N; ++j) cand[j] = fmin(fmax(th[j] + sc * d[j], prm.lo[j]), prm.hi[j]);
                const double e = ik_pose_error<T, N>(rb, cand, Td);
                if (e < bst_err) {
                    bst_err = e;
#pragma unroll
                    for (int j = 0; j < N; ++j) bst[j] = cand[j];
                }
            }
            if (bst_err < cur * 1.1) {
#pragma unroll
                for (int j = 0; j < N; ++j) th[j] = bst[j];
            } else {
#pragma unroll
                for (int j = 0; j < N; ++j) th[j] = fmin(fmax(th[j] + T(0.1) * d[j], prm.lo[j]), prm.hi[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < N; ++j) th[j] = fmin(fmax(th[j] + scale * d[j], prm.lo[j]), prm.hi[j]);
        }
    }
    st.k = k;
    if (!ok && k < prm.max_iter) return false;  // window exhausted, budget not
    if (!ok && st.best_err < cur) {
        // max_iterations reached (ik.py:264-275): fall back to the best iterate if it is better
#pragma unroll
        for (int j = 0; j < N; ++j) th[j] = st.best[j];
        JointCS<T, N> q;
        joint_cs(rb, th, q);
        double Tc[16], V[6];
        fk_jacobian<T, N>(rb, q, Tc, (T *)nullptr);
        ik_error(Tc, Td, V, rot, trans);
        ok = rot < prm.eomg && trans < prm.ev;
    }
    iterations = k + 1;
    return true;
}

// One target, whole iteration budget.  th: initial guess in, solution out.
template <typename T, int N>
MPK_HD bool ik_dls(const RobotPack<T, N> &rb, const double *Td, T (&th)[N], const IkParams<T, MPK_MAX_DOF_> &prm,
                   unsigned long long seed, unsigned long long target, T *J, int &iterations,
                   const double *noise = nullptr, int noise_rows = 0, int *restarts = nullptr) {
    IkState<T, N> st;
    ik_state_init(st, th, prm);
    bool ok;
    ik_dls_window(rb, Td, st, prm, seed, target, J, prm.max_iter, ok, iterations, noise, noise_rows);
    if (restarts) *restarts = st.restarts;
#pragma unroll
    for (int j = 0; j < N; ++j) th[j] = st.th[j];
    return ok;
}

// ---- time scaling (planning/trajectory.py:15-75) -------------------------------
// float64, the reference's operation order, no FMA contraction (explicit _rn intrinsics),
// so that the single rounding to float32 reproduces the reference bit for bit.
struct TimeScale {
    double s, sd, sdd;
};

// `method`: 3 cubic, 5 quintic.  Anything else gives zero scaling (the planner's CPU kernel,
// planning/trajectory.py:67-68) -- unless MPK_TRAJ_REGISTRY_CONTRACT (0x100) is or-ed in, which
// selects the contract of the registry launchers and their kernels
// (cuda_kernels/trajectory_kernels.py:40-76, 179, 195-198): linear scaling for any other method,
// and s = ds = dds = 0 ("sit at start") when N <= 1 or Tf <= 0.
MPK_HD TimeScale time_scaling(int64_t idx, int64_t N, double Tf, int method) {
    const bool registry = (method & 0x100) != 0;
    method &= 0xff;
    if (registry && (N <= 1 || !(Tf > 0.0))) return TimeScale{0.0, 0.0, 0.0};
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
    } else if (registry) {
        r.s = tau;
        r.sd = rn_div(1.0, Tf);
        r.sdd = 0.0;
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
