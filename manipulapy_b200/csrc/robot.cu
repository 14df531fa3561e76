// robot.cu -- host side of the C ABI: robot constant pack construction, error
// reporting, and the FMA peak micro-benchmark.
//
// mpk_robot_create re-expresses the reference's constant pack (S_list, M, Glist,
// Mlist_per_link; dynamics/manipulator_dynamics.py:46-75) in joint-aligned link
// frames (see mpk_device.cuh).  Joint i with space screw S_i = (w, v) at the home
// configuration defines a line:
//   revolute (|w| = 1): direction z = w through q = w x v (the point of the axis closest to
//       the space origin); the pitch w.v must be 0 (helical joints are rejected);
//   prismatic (w = 0):  direction z = v/|v|, st = |v|; the line's position is free and is
//       put through the previous axis.
// Frame i has its z axis on line i.  Its x axis and origin are chosen from the NEXT line the
// way Denavit and Hartenberg do (x_i along the common normal of lines i and i+1, origin at
// its foot), or, when the two lines are nearly parallel (|z_i x z_{i+1}| < 0.1, where the
// common normal is ill-conditioned or undefined), the way Hayati does (origin kept where the
// previous normal met line i, x_i towards the point where line i+1 pierces the plane through
// that origin normal to z_i).  Either way the home pose of frame i in frame i-1 is exactly
//   X_i = Tx(a_i) Rx(alpha_i) Ry(beta_i) Rz(phi_i) Tz(d_i),   beta_i = 0 in the D-H case,
// and with F_i the home pose of frame i in space,  e^{[S_i] th} F_i = F_i Jz(th), so
//   prod_j e^{[S_j] th_j} = F_0 Jz(th_0) prod_{j>=1} (X_j Jz(th_j)) F_{n-1}^{-1},
// which is the identity the kernels rely on.  The factorisation is verified against
// F_{i-1}^{-1} F_i before the pack is accepted.
#include <cmath>
#include <cstring>

#include "mpk_common.cuh"

namespace mpk {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MPK_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return MPK_OK;
}

namespace {

struct SE3 {
    double R[9];
    double p[3];
};

SE3 from_mat4(const double *T) {
    SE3 o;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) o.R[3 * r + c] = T[4 * r + c];
        o.p[r] = T[4 * r + 3];
    }
    return o;
}
SE3 identity() {
    SE3 o;
    std::memset(&o, 0, sizeof o);
    o.R[0] = o.R[4] = o.R[8] = 1.0;
    return o;
}
SE3 inverse(const SE3 &a) {
    SE3 o;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) o.R[3 * r + c] = a.R[3 * c + r];
    for (int r = 0; r < 3; ++r)
        o.p[r] = -(o.R[3 * r] * a.p[0] + o.R[3 * r + 1] * a.p[1] + o.R[3 * r + 2] * a.p[2]);
    return o;
}
SE3 mul(const SE3 &a, const SE3 &b) {
    SE3 o;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
            o.R[3 * r + c] =
                a.R[3 * r] * b.R[c] + a.R[3 * r + 1] * b.R[3 + c] + a.R[3 * r + 2] * b.R[6 + c];
        o.p[r] = a.p[r] + a.R[3 * r] * b.p[0] + a.R[3 * r + 1] * b.p[1] + a.R[3 * r + 2] * b.p[2];
    }
    return o;
}
// 6x6 adjoint [[R,0],[[p]R,R]] (twists [w; v])
void adjoint(const SE3 &T, double *Ad) {
    const double *p = T.p;
    const double P[9] = {0, -p[2], p[1], p[2], 0, -p[0], -p[1], p[0], 0};
    std::memset(Ad, 0, 36 * sizeof(double));
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += P[3 * r + k] * T.R[3 * k + c];
            Ad[6 * r + c] = T.R[3 * r + c];
            Ad[6 * (r + 3) + c + 3] = T.R[3 * r + c];
            Ad[6 * (r + 3) + c] = s;
        }
}
void cross(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
double norm3(const double *a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// Right-handed frame with the given unit z; x is the world axis least aligned with z,
// orthogonalised.  R columns = (x, y, z).
void frame_from_z(const double *z, double *R) {
    int k = 0;
    if (std::fabs(z[1]) < std::fabs(z[k])) k = 1;
    if (std::fabs(z[2]) < std::fabs(z[k])) k = 2;
    double e[3] = {0, 0, 0};
    e[k] = 1.0;
    const double d = z[k];
    double x[3] = {e[0] - d * z[0], e[1] - d * z[1], e[2] - d * z[2]};
    const double nx = norm3(x);
    for (double &c : x) c /= nx;
    double y[3];
    cross(z, x, y);
    for (int r = 0; r < 3; ++r) {
        R[3 * r] = x[r];
        R[3 * r + 1] = y[r];
        R[3 * r + 2] = z[r];
    }
}

double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

SE3 frame_from_axes(const double *x, const double *z, const double *o) {
    SE3 F;
    double y[3];
    cross(z, x, y);
    for (int r = 0; r < 3; ++r) {
        F.R[3 * r] = x[r];
        F.R[3 * r + 1] = y[r];
        F.R[3 * r + 2] = z[r];
        F.p[r] = o[r];
    }
    return F;
}

SE3 rot_x(double c, double s) {
    SE3 o = identity();
    o.R[4] = c; o.R[5] = -s; o.R[7] = s; o.R[8] = c;
    return o;
}
SE3 rot_y(double c, double s) {
    SE3 o = identity();
    o.R[0] = c; o.R[2] = s; o.R[6] = -s; o.R[8] = c;
    return o;
}
SE3 rot_z(double c, double s) {
    SE3 o = identity();
    o.R[0] = c; o.R[1] = -s; o.R[3] = s; o.R[4] = c;
    return o;
}
SE3 trans(double x, double y, double z) {
    SE3 o = identity();
    o.p[0] = x; o.p[1] = y; o.p[2] = z;
    return o;
}

// One joint axis at the home configuration.
struct Line {
    double z[3], q[3];
    bool prismatic;
};

// Link-frame construction (see the header comment).  F[i]: home pose of frame i in space.
// Fills a, ca, sa, cb, sb, phi, d (and cphi, sphi) of the pack; returns the largest deviation
// between the factored X_i and F_{i-1}^{-1} F_i.
double build_frames(int n, Line *L, RobotPack<double, MPK_MAX_DOF> &pk, SE3 *F) {
    double xin[3], r[3];  // x axis and point of line i handed over by the previous pair
    {
        double R0[9];
        frame_from_z(L[0].z, R0);
        for (int k = 0; k < 3; ++k) {
            xin[k] = R0[3 * k];
            if (L[0].prismatic) L[0].q[k] = 0.0;
            r[k] = L[0].q[k];
        }
    }
    double worst = 0.0;
    SE3 Fprev = identity();
    for (int i = 0; i < n; ++i) {
        double x[3], o[3], xin_next[3] = {0, 0, 0}, r_next[3] = {0, 0, 0};
        double a = 0, ca = 1, sa = 0, cb = 1, sb = 0;
        if (i + 1 < n) {
            Line &A = L[i], &Bn = L[i + 1];
            if (Bn.prismatic)
                for (int k = 0; k < 3; ++k) Bn.q[k] = r[k];
            double cr[3];
            cross(A.z, Bn.z, cr);
            const double sn = norm3(cr), cd = dot3(A.z, Bn.z);
            if (sn >= 0.1) {
                // Denavit-Hartenberg: x along the common normal, origin at its foot on line i
                for (int k = 0; k < 3; ++k) x[k] = cr[k] / sn;
                double dq[3] = {Bn.q[0] - A.q[0], Bn.q[1] - A.q[1], Bn.q[2] - A.q[2]};
                a = dot3(dq, x);
                const double d1 = dot3(dq, A.z), d2 = dot3(dq, Bn.z), den = 1.0 - cd * cd;
                const double t1 = (d1 - cd * d2) / den, t2 = (cd * d1 - d2) / den;
                for (int k = 0; k < 3; ++k) {
                    o[k] = A.q[k] + t1 * A.z[k];
                    r_next[k] = Bn.q[k] + t2 * Bn.z[k];
                    xin_next[k] = x[k];
                }
                const double nrm = std::sqrt(sn * sn + cd * cd);
                sa = sn / nrm;
                ca = cd / nrm;
            } else {
                // Hayati: origin stays at the handed-over point; x towards the point where line
                // i+1 pierces the plane through it normal to z_i
                for (int k = 0; k < 3; ++k) o[k] = r[k];
                double oq[3] = {o[0] - Bn.q[0], o[1] - Bn.q[1], o[2] - Bn.q[2]};
                const double t = dot3(oq, A.z) / cd;
                double perp[3];
                for (int k = 0; k < 3; ++k) {
                    r_next[k] = Bn.q[k] + t * Bn.z[k];
                    perp[k] = r_next[k] - o[k];
                }
                const double pz = dot3(perp, A.z);
                for (int k = 0; k < 3; ++k) perp[k] -= pz * A.z[k];
                a = norm3(perp);
                if (a > 1e-12) {
                    for (int k = 0; k < 3; ++k) x[k] = perp[k] / a;
                } else {
                    a = 0.0;
                    for (int k = 0; k < 3; ++k) x[k] = xin[k];
                }
                double y[3];
                cross(A.z, x, y);
                double zx = dot3(Bn.z, x);
                const double zy = dot3(Bn.z, y), zz = dot3(Bn.z, A.z);
                if (std::fabs(zx) < 1e-14) zx = 0.0;
                sb = zx;
                cb = std::sqrt(1.0 - sb * sb);
                const double nrm = std::sqrt(zy * zy + zz * zz);
                sa = -zy / nrm;
                ca = zz / nrm;
                if (std::fabs(sa) < 1e-14) {  // exactly parallel (or anti-parallel) lines
                    sa = 0.0;
                    ca = ca > 0 ? 1.0 : -1.0;
                }
                // x axis handed to frame i+1: Rx(alpha) Ry(beta) e_x in frame-i coordinates
                const double hx = cb, hy = sa * sb, hz = -ca * sb;
                for (int k = 0; k < 3; ++k) xin_next[k] = hx * x[k] + hy * y[k] + hz * A.z[k];
            }
        } else {
            for (int k = 0; k < 3; ++k) {
                x[k] = xin[k];
                o[k] = r[k];
            }
        }
        // re-orthogonalise x against z (rounding) and build the frame
        {
            const double xz = dot3(x, L[i].z);
            for (int k = 0; k < 3; ++k) x[k] -= xz * L[i].z[k];
            const double nx = norm3(x);
            for (int k = 0; k < 3; ++k) x[k] /= nx;
        }
        F[i] = frame_from_axes(x, L[i].z, o);
        // offsets of this frame against what the previous pair handed over
        double phi = 0.0, d = 0.0;
        if (i > 0) {
            double cx[3];
            cross(xin, x, cx);
            phi = std::atan2(dot3(cx, L[i].z), dot3(xin, x));
            double ro[3] = {o[0] - r[0], o[1] - r[1], o[2] - r[2]};
            d = dot3(ro, L[i].z);
        }
        pk.phi[i] = phi;
        pk.d[i] = d;
        pk.cphi[i] = std::cos(phi);
        pk.sphi[i] = std::sin(phi);
        if (i == 0) {
            pk.a[0] = 0; pk.ca[0] = 1; pk.sa[0] = 0; pk.cb[0] = 1; pk.sb[0] = 0;
            for (int k = 0; k < 9; ++k) pk.Rb[k] = F[0].R[k];
            for (int k = 0; k < 3; ++k) pk.pb[k] = F[0].p[k];
        } else {
            // check the factorisation against the geometric relative pose
            const SE3 X = mul(inverse(Fprev), F[i]);
            const SE3 Y = mul(mul(mul(trans(pk.a[i], 0, 0), rot_x(pk.ca[i], pk.sa[i])),
                                  mul(rot_y(pk.cb[i], pk.sb[i]), rot_z(pk.cphi[i], pk.sphi[i]))),
                              trans(0, 0, d));
            for (int k = 0; k < 9; ++k) worst = std::fmax(worst, std::fabs(X.R[k] - Y.R[k]));
            for (int k = 0; k < 3; ++k) worst = std::fmax(worst, std::fabs(X.p[k] - Y.p[k]));
        }
        if (i + 1 < n) {
            pk.a[i + 1] = a;
            pk.ca[i + 1] = ca;
            pk.sa[i + 1] = sa;
            pk.cb[i + 1] = cb;
            pk.sb[i + 1] = sb;
        }
        for (int k = 0; k < 3; ++k) {
            xin[k] = xin_next[k];
            r[k] = r_next[k];
        }
        Fprev = F[i];
    }
    return worst;
}

}  // namespace
}  // namespace mpk

using namespace mpk;

extern "C" int mpk_version(void) { return 110; }  // 110: + mpk_inverse_kinematics_dls_modes
extern "C" const char *mpk_last_error(void) { return g_err.c_str(); }

extern "C" int mpk_robot_create(int n, const double *S_list, const double *M, const double *Glist,
                                const double *Mcom, int flags, mpk_robot **out) {
    if (!out) return fail(MPK_EINVAL, "out is NULL");
    *out = nullptr;
    if (n < 1 || n > MPK_MAX_DOF)
        return fail(MPK_EUNSUPPORTED, "supported joint counts are 1.." + std::to_string(MPK_MAX_DOF));
    if (!S_list || !M) return fail(MPK_EINVAL, "S_list and M are required");
    if ((Glist == nullptr) != (Mcom == nullptr))
        return fail(MPK_EINVAL, "Glist and Mlist_per_link must be given together");

    mpk_robot *rb = new mpk_robot();
    std::memset(rb, 0, sizeof *rb);
    rb->n = n;
    rb->has_dynamics = Glist != nullptr;
    fill_trig_table(rb->pack.trig);
    bool rigid = true;

    Line lines[MPK_MAX_DOF];
    bool all_rev = true;
    for (int i = 0; i < n; ++i) {
        double w[3], v[3];
        for (int r = 0; r < 3; ++r) {
            w[r] = S_list[r * n + i];
            v[r] = S_list[(r + 3) * n + i];
        }
        Line &ln = lines[i];
        const double nw = norm3(w), nv = norm3(v);
        if (nw == 0.0) {
            all_rev = false;
            ln.prismatic = true;
            rb->pack.sr[i] = 0.0;
            rb->pack.st[i] = nv;
            ln.z[0] = 0; ln.z[1] = 0; ln.z[2] = 1;
            if (nv > 0.0)
                for (int r = 0; r < 3; ++r) ln.z[r] = v[r] / nv;
            ln.q[0] = ln.q[1] = ln.q[2] = 0.0;  // placed by build_frames
        } else {
            if (std::fabs(nw - 1.0) > 1e-9) {
                delete rb;
                return fail(MPK_EUNSUPPORTED,
                            "screw axis " + std::to_string(i) +
                                " has a non-unit angular part; the reference's transform_from_twist "
                                "(utils/se3.py:33-42) is only a rigid motion for unit omega or omega = 0");
            }
            ln.prismatic = false;
            for (int r = 0; r < 3; ++r) ln.z[r] = w[r] / nw;
            cross(ln.z, v, ln.q);  // q = w x v
            const double h = dot3(ln.z, v);
            if (std::fabs(h) > 1e-12 * (nv > 1.0 ? nv : 1.0)) {
                delete rb;
                return fail(MPK_EUNSUPPORTED,
                            "screw axis " + std::to_string(i) +
                                " is helical (omega . v != 0); only revolute (v = -omega x q) and "
                                "prismatic (omega = 0) joints are supported");
            }
            rb->pack.sr[i] = 1.0;
            rb->pack.st[i] = 0.0;
        }
    }
    SE3 frames[MPK_MAX_DOF];
    const double dev = build_frames(n, lines, rb->pack, frames);
    if (!(dev < 1e-10)) {
        delete rb;
        return fail(MPK_EUNSUPPORTED, "link-frame factorisation failed (deviation " + std::to_string(dev) + ")");
    }
    rb->all_revolute = all_rev ? 1 : 0;
    rb->plain = rb->all_revolute;
    rb->first_revolute = rb->pack.sr[0] != 0.0;
    for (int i = 0; i < n; ++i)
        if (rb->pack.sb[i] != 0.0) rb->plain = 0;
    // Offsets that are pure rounding noise of the frame construction (|x| < 1e-15 m, e.g. the 1e-17
    // common-normal length of two axes that intersect) are exact zeros; then classify the links
    // (mpk_device.cuh "link geometry classes").
    rb->geo = 0;
    for (int i = 1; i < n; ++i) {
        if (std::fabs(rb->pack.a[i]) < 1e-15) rb->pack.a[i] = 0.0;
        if (std::fabs(rb->pack.d[i]) < 1e-15 && rb->pack.st[i] == 0.0) rb->pack.d[i] = 0.0;
        unsigned cls = 0;
        if (rb->pack.sa[i] == 1.0) cls |= kGeoPerp;
        else if (rb->pack.sa[i] == 0.0 && rb->pack.ca[i] == 1.0) cls |= kGeoPar;
        if (rb->pack.a[i] == 0.0) cls |= kGeoA0;
        if (rb->pack.d[i] == 0.0 && rb->pack.st[i] == 0.0) cls |= kGeoD0;
        rb->geo |= cls << (4 * i);
    }
    if (!rb->plain) rb->geo = 0;

    for (int i = 0; i < n; ++i) {
        const SE3 &F = frames[i];
        if (Glist) {
            const double *G = Glist + 36 * i;
            double Gs[36];
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 6; ++c) Gs[6 * r + c] = 0.5 * (G[6 * r + c] + G[6 * c + r]);
            // rigid body at its centre of mass: block diagonal [I, m 1]
            bool rg = Gs[21] == Gs[28] && Gs[21] == Gs[35];
            for (int r = 0; r < 3 && rg; ++r)
                for (int c = 0; c < 3; ++c) {
                    if (Gs[6 * r + c + 3] != 0.0) rg = false;
                    if (r != c && Gs[6 * (r + 3) + c + 3] != 0.0) rg = false;
                }
            rigid = rigid && rg;
            const SE3 Mc = from_mat4(Mcom + 16 * i);
            const SE3 C = mul(inverse(Mc), F);   // V_com = Ad(C) V_frame
            const SE3 Ci = mul(inverse(F), Mc);  // CoM frame pose in frame i
            double Ad[36], GA[36], Gp[36];
            adjoint(C, Ad);
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 6; ++c) {
                    double s = 0;
                    for (int k = 0; k < 6; ++k) s += Gs[6 * r + k] * Ad[6 * k + c];
                    GA[6 * r + c] = s;
                }
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 6; ++c) {
                    double s = 0;
                    for (int k = 0; k < 6; ++k) s += Ad[6 * k + r] * GA[6 * k + c];
                    Gp[6 * r + c] = s;
                }
            int t = 0;
            for (int r = 0; r < 6; ++r)
                for (int c = r; c < 6; ++c) rb->pack.G[i][t++] = 0.5 * (Gp[6 * r + c] + Gp[6 * c + r]);
            for (int k = 0; k < 3; ++k) rb->pack.cg[i][k] = Ci.p[k];
            rb->pack.mg[i] = G[21];
            if (rg) {
                // analytic rigid form: I about the frame origin = Rc Icom Rc^T + m((c.c)1 - c c^T)
                const double m = Gs[21];
                const double *c = Ci.p;
                double RI[9], Ir[9];
                for (int r = 0; r < 3; ++r)
                    for (int cc = 0; cc < 3; ++cc) {
                        double s = 0;
                        for (int k = 0; k < 3; ++k) s += Ci.R[3 * r + k] * Gs[6 * k + cc];
                        RI[3 * r + cc] = s;
                    }
                for (int r = 0; r < 3; ++r)
                    for (int cc = 0; cc < 3; ++cc) {
                        double s = 0;
                        for (int k = 0; k < 3; ++k) s += RI[3 * r + k] * Ci.R[3 * cc + k];
                        Ir[3 * r + cc] = s;
                    }
                rb->pack.Ic[i][0] = Ir[0];
                rb->pack.Ic[i][1] = 0.5 * (Ir[1] + Ir[3]);
                rb->pack.Ic[i][2] = 0.5 * (Ir[2] + Ir[6]);
                rb->pack.Ic[i][3] = Ir[4];
                rb->pack.Ic[i][4] = 0.5 * (Ir[5] + Ir[7]);
                rb->pack.Ic[i][5] = Ir[8];
                for (int k = 0; k < 3; ++k) rb->pack.com[i][k] = c[k];
                const double c2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
                for (int r = 0; r < 3; ++r)
                    for (int cc = 0; cc < 3; ++cc)
                        Ir[3 * r + cc] += m * ((r == cc ? c2 : 0.0) - c[r] * c[cc]);
                rb->pack.I[i][0] = Ir[0];
                rb->pack.I[i][1] = 0.5 * (Ir[1] + Ir[3]);
                rb->pack.I[i][2] = 0.5 * (Ir[2] + Ir[6]);
                rb->pack.I[i][3] = Ir[4];
                rb->pack.I[i][4] = 0.5 * (Ir[5] + Ir[7]);
                rb->pack.I[i][5] = Ir[8];
                for (int k = 0; k < 3; ++k) rb->pack.h[i][k] = m * c[k];
                rb->pack.m[i] = m;
            }
        }
    }
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 9; ++k) rb->F[i][k] = frames[i].R[k];
        for (int k = 0; k < 3; ++k) rb->F[i][9 + k] = frames[i].p[k];
    }
    const SE3 E = mul(inverse(frames[n - 1]), from_mat4(M));
    for (int k = 0; k < 9; ++k) rb->pack.Ree[k] = E.R[k];
    for (int k = 0; k < 3; ++k) rb->pack.pee[k] = E.p[k];
    rb->rigid = (rigid && !(flags & MPK_ROBOT_FORCE_GENERAL)) ? 1 : 0;
    *out = rb;
    return MPK_OK;
}

extern "C" unsigned mpk_robot_geometry_signature(const mpk_robot *rb) { return rb ? rb->geo : 0u; }

extern "C" int mpk_robot_link_geometry(const mpk_robot *rb, double *out) {
    if (!rb || !out) return fail(MPK_EINVAL, "robot / out is NULL");
    for (int i = 0; i < rb->n; ++i) {
        const auto &p = rb->pack;
        const double row[8] = {p.a[i], p.ca[i], p.sa[i], p.cb[i], p.sb[i], p.phi[i], p.d[i], p.sr[i]};
        for (int k = 0; k < 8; ++k) out[8 * i + k] = row[k];
    }
    return MPK_OK;
}

extern "C" void mpk_robot_destroy(mpk_robot *rb) { delete rb; }
extern "C" int mpk_robot_dof(const mpk_robot *rb) { return rb ? rb->n : MPK_EINVAL; }
extern "C" int mpk_robot_is_rigid(const mpk_robot *rb) { return rb ? rb->rigid : MPK_EINVAL; }
extern "C" int mpk_robot_all_revolute(const mpk_robot *rb) { return rb ? rb->all_revolute : MPK_EINVAL; }

// ---- FMA peak micro-benchmark ---------------------------------------------------
// mode 0: 8 dependent chains a = a*m + b per thread, m and b shared (operand reuse: the
//         datasheet-style peak);
// mode 1: 12 registers rotating, x_k = fma(x_{k+1}, x_{k+2}, x_{k+3}): three DISTINCT register
//         operands per instruction, like the rigid-body algebra of the kernels;
// mode 2: mode 1 with one operand taken from the constant bank (a kernel parameter);
// mode 3: the planar-rotation pattern of the kernels: DMUL + DFMA pairs on distinct registers
//         (two dependent rotations per iteration; every operand is rewritten in the loop).
// All modes execute 16 flops per loop iteration per chain slot so the callers' flop count
// (blocks * threads * iters * 16) holds: modes 1-3 run 8 instructions per iteration too.
struct PeakConsts {
    double k[8];
};

template <typename T, int MODE>
__global__ void fma_peak_kernel(int64_t iters, double *sink, const __grid_constant__ PeakConsts pc) {
    if (MODE == 0) {
        T a0 = T(threadIdx.x) * T(1e-3), a1 = a0 + T(1), a2 = a0 + T(2), a3 = a0 + T(3);
        T a4 = a0 + T(4), a5 = a0 + T(5), a6 = a0 + T(6), a7 = a0 + T(7);
        const T m = T(0.999999), b = T(1e-7);
        for (int64_t i = 0; i < iters; ++i) {
            a0 = a0 * m + b; a1 = a1 * m + b; a2 = a2 * m + b; a3 = a3 * m + b;
            a4 = a4 * m + b; a5 = a5 * m + b; a6 = a6 * m + b; a7 = a7 * m + b;
        }
        const T s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
        if (s == T(-1.2345)) sink[0] = (double)s;  // never true; keeps the chains alive
    } else {
        T x[12];
        unsigned iq[4] = {threadIdx.x, threadIdx.x * 3u + 1u, threadIdx.x * 5u + 2u, threadIdx.x * 7u + 3u};
#pragma unroll
        for (int k = 0; k < 12; ++k) x[k] = T(0.5) + T(threadIdx.x + k) * T(1e-4);
        for (int64_t i = 0; i < iters; ++i) {
            if (MODE == 1) {
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = x[k + 1] * x[k + 2] - x[k + 3];
            } else if (MODE == 2) {
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = x[k + 1] * T(pc.k[k]) - x[k + 3];
            } else if (MODE == 4) {
                // half of the instructions with a uniform-register operand, half with three registers
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    x[k] = (k & 1) ? x[k + 1] * x[k + 2] - x[k + 3] : x[k + 1] * T(pc.k[k]) - x[k + 3];
            } else if (MODE == 5) {
                // three registers, but consecutive instructions share their middle operand (.reuse)
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = x[k + 1] * x[(k < 4) ? 10 : 11] - x[k + 3];
            } else if (MODE == 7 || MODE == 8) {
                // each DFMA (7: three registers; 8: one constant-bank operand) followed by one 32-bit
                // integer multiply-add on three registers: does instruction issue / register-file
                // bandwidth taken by the non-fp64 half of a kernel slow the fp64 pipe down?
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    x[k] = MODE == 7 ? x[k + 1] * x[k + 2] - x[k + 3] : x[k + 1] * T(pc.k[k]) - x[k + 3];
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(iq[k & 3]) : "r"(iq[(k + 1) & 3]), "r"(iq[(k + 2) & 3]));
                }
            } else if (MODE == 6) {
                // three registers, consecutive instructions share TWO operands pairwise (a*b - c, a*b' - c)
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = x[(k & ~1) + 1] * x[k + 2] - x[(k & ~1) + 3];
            } else {
                // the planar-rotation pattern of the kernels, p' = c p - s q, q' = s p + c q: two DMUL +
                // two DFMA on distinct registers per rotation, two rotations per iteration (8
                // instructions, 12 flops -- counted as 16).  Every register the loop reads is also
                // written by it, so nothing is loop-invariant.  (The round-1 version of this mode left
                // x[1], x[3], x[5], x[7] unwritten: ptxas hoisted two of its four DMULs out of the loop,
                // the loop body had 6 instructions where the caller counted 8, and the reported rate came
                // out 8/6 above the pipe's peak -- profiles/r2_fp64_operand_patterns.md.)
                {
                    const T p = x[0], q = x[1], c = x[2], sn = x[3];
                    const T t0 = sn * q, t1 = sn * p;
                    x[0] = c * p - t0;
                    x[1] = c * q + t1;
                }
                {
                    const T p = x[2], q = x[3], c = x[0], sn = x[1];
                    const T t0 = sn * q, t1 = sn * p;
                    x[2] = c * p - t0;
                    x[3] = c * q + t1;
                }
                // keep the values bounded (the rotation scales by |(c, s)|): renormalising would add
                // instructions, so the chains are re-seeded from the untouched registers instead
                if ((i & 63) == 63) {
                    x[0] = x[4]; x[1] = x[5]; x[2] = x[6]; x[3] = x[7];
                }
                continue;
            }
            // rotate so that the next iteration's operands are other registers
            const T t0 = x[8];
            x[8] = x[9]; x[9] = x[10]; x[10] = x[11]; x[11] = x[0]; x[0] = t0;
        }
        T s = T(0);
#pragma unroll
        for (int k = 0; k < 12; ++k) s += x[k];
        if (s == T(-1.2345) || (iq[0] ^ iq[1] ^ iq[2] ^ iq[3]) == 0x12345u) sink[0] = (double)s;
    }
}

// ---- store-bandwidth micro-benchmark ---------------------------------------------------
// The write-only ceiling of this GPU's HBM for the roofline of the store-bound kernels (trajectory
// rows): every thread writes 16-byte vectors, a warp 512 contiguous bytes per instruction.
//   mode 0: st.global (default, write-back)   mode 1: st.global.cs (streaming, what the kernels use)
//   mode 2: st.global.cg                      mode 3: st.global.wt
template <int MODE>
__global__ void __launch_bounds__(256) store_peak_kernel(float4 *dst, int64_t n16) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float4 v = make_float4(1.f, 2.f, 3.f, (float)blockIdx.x);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        if (MODE == 0) dst[i] = v;
        else if (MODE == 1) __stcs(dst + i, v);
        else if (MODE == 2) __stcg(dst + i, v);
        else __stwt(dst + i, v);
    }
}

extern "C" int mpk_fma_peak(int dtype, int blocks, int threads, int64_t iters, double *sink_dev,
                            void *stream) {
    if (blocks <= 0 || threads <= 0 || threads > 1024 || iters < 0 || !sink_dev)
        return fail(MPK_EINVAL, "bad fma_peak arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int mode = dtype >> 8;
    dtype &= 0xff;
    PeakConsts pc;
    for (int k = 0; k < 8; ++k) pc.k[k] = 0.75 + 0.03 * k;
    if (dtype == MPK_F64 && mode == 0) fma_peak_kernel<double, 0><<<blocks, threads, 0, s>>>(iters, sink_dev, pc);
    else if (dtype == MPK_F64 && mode == 1) fma_peak_kernel<double, 1><<<blocks, threads, 0, s>>>(iters, sink_dev, pc);
    else if (dtype == MPK_F64 && mode == 2) fma_peak_kernel<double, 2><<<blocks, threads, 0, s>>>(iters, sink_dev, pc);
    else if (dtype == MPK_F64 && mode == 3) fma_peak_kernel<double, 3><<<blocks, threads, 0, s>>>(iters, sink_dev, pc);
    else if (dtype == MPK_F64 && mode == 4) fma_peak_kernel<double, 4><<<blocks, threads, 0, s>>>(iters, sink_dev, pc);
    else if (dtype == MPK_F64 && mode == 5) fma_peak_kernel<double, 5><<<blocks, threads, 0, s>>>(iters, sink_dev, pc);
    else if (dtype == MPK_F64 && mode == 6) fma_peak_kernel<double, 6><<<blocks, threads, 0, s>>>(iters, sink_dev, pc);
    else if (dtype == MPK_F64 && mode == 7) fma_peak_kernel<double, 7><<<blocks, threads, 0, s>>>(iters, sink_dev, pc);
    else if (dtype == MPK_F64 && mode == 8) fma_peak_kernel<double, 8><<<blocks, threads, 0, s>>>(iters, sink_dev, pc);
    else if (dtype == MPK_F32 && mode == 0) fma_peak_kernel<float, 0><<<blocks, threads, 0, s>>>(iters, sink_dev, pc);
    else if (dtype == MPK_F32 && mode == 1) fma_peak_kernel<float, 1><<<blocks, threads, 0, s>>>(iters, sink_dev, pc);
    else return fail(MPK_EINVAL, "bad dtype / mode");
    return check_launch("fma_peak");
}

extern "C" int mpk_store_peak(void *dst_dev, int64_t bytes, int mode, int blocks, void *stream) {
    if (!dst_dev || bytes < 16 || !aligned16(dst_dev) || blocks <= 0 || mode < 0 || mode > 3)
        return fail(MPK_EINVAL, "bad store_peak arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float4 *d = static_cast<float4 *>(dst_dev);
    const int64_t n16 = bytes / 16;
    if (mode == 0) store_peak_kernel<0><<<blocks, 256, 0, s>>>(d, n16);
    else if (mode == 1) store_peak_kernel<1><<<blocks, 256, 0, s>>>(d, n16);
    else if (mode == 2) store_peak_kernel<2><<<blocks, 256, 0, s>>>(d, n16);
    else store_peak_kernel<3><<<blocks, 256, 0, s>>>(d, n16);
    return check_launch("store_peak");
}
