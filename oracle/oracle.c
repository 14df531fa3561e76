/*
 * oracle.c -- CPU restatement (plain C, float64) of ManipulaPy v1.4.1's
 * batched trajectory-and-dynamics hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the
 * __graft_entry__.smoke() check and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product path (manipulapy_b200/) never calls it and has
 * no CPU fallback.
 *
 * Every function follows the reference's *own* algorithm (same order of
 * operations, same finite-difference Coriolis, same clipping), citing the
 * reference file:line it restates (paths relative to the reference checkout).
 * The pinning of this oracle against the reference's golden vectors
 * (tests/data/dynamics_golden_{ur5,panda}.npz) and against outputs of the
 * reference run in the build container lives in tests/test_oracle_golden.py.
 *
 * Two families:
 *   orc_*            literal restatement (mass matrix = sum_k J_k^T G_k J_k,
 *                    Coriolis via central finite differences with eps = 1e-6).
 *   orc_*_analytic   the same quantities through a Modern-Robotics body-frame
 *                    Newton-Euler recursion in the link-CoM frames (SURVEY.md
 *                    App. C).  No finite-difference noise; used for long
 *                    rollouts / large samples and as the fast CPU port.
 *
 * Conventions (reference): twists [omega; v], wrenches [moment; force],
 * S_list is (6, n) row-major (column per joint), matrices row-major.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ORC_MAX_DOF 16

typedef struct {
    int n;
    const double *S;    /* (6, n) row-major: S[r*n + j]                      */
    const double *M;    /* (4, 4) end-effector home pose                       */
    const double *G;    /* (n, 6, 6) spatial inertias in the link-CoM frames   */
    const double *Mcom; /* (n, 4, 4) link-CoM home poses (Mlist_per_link)      */
} orc_robot;

/* ------------------------------------------------------------------ small algebra */

static void mat4_mul(const double *A, const double *B, double *C) {
    double t[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j];
            t[i * 4 + j] = s;
        }
    memcpy(C, t, sizeof t);
}

static void mat4_eye(double *T) {
    memset(T, 0, 16 * sizeof(double));
    T[0] = T[5] = T[10] = T[15] = 1.0;
}

/* utils/se3.py:95-100 TransInv: [R^T, -R^T p] (the reference's np.linalg.inv of
 * an SE(3) matrix equals this up to rounding). */
static void se3_inv(const double *T, double *Ti) {
    double t[16];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) t[i * 4 + j] = T[j * 4 + i];
    for (int i = 0; i < 3; ++i) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += t[i * 4 + k] * T[k * 4 + 3];
        t[i * 4 + 3] = -s;
    }
    t[12] = t[13] = t[14] = 0.0;
    t[15] = 1.0;
    memcpy(Ti, t, sizeof t);
}

/* utils/se3.py:33-42 transform_from_twist: Rodrigues with the screw used as
 * given (unit omega, or omega = 0 for a prismatic joint -- never normalised). */
void orc_transform_from_twist(const double *S6, double theta, double *T) {
    const double w0 = S6[0], w1 = S6[1], w2 = S6[2];
    const double W[9] = {0.0, -w2, w1, w2, 0.0, -w0, -w1, w0, 0.0};
    double W2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += W[i * 3 + k] * W[k * 3 + j];
            W2[i * 3 + j] = s;
        }
    const double st = sin(theta), ct = cos(theta);
    double R[9], Gm[9];
    for (int i = 0; i < 9; ++i) {
        const double I = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = I + st * W[i] + (1.0 - ct) * W2[i];
        Gm[i] = I * theta + (1.0 - ct) * W[i] + (theta - st) * W2[i];
    }
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[i * 4 + j] = R[i * 3 + j];
        T[i * 4 + 3] = Gm[i * 3 + 0] * S6[3] + Gm[i * 3 + 1] * S6[4] + Gm[i * 3 + 2] * S6[5];
    }
    T[12] = T[13] = T[14] = 0.0;
    T[15] = 1.0;
}

/* utils/se3.py:45-52 adjoint_transform: [[R, 0], [[p]R, R]]. */
void orc_adjoint(const double *T, double *Ad) {
    const double p0 = T[3], p1 = T[7], p2 = T[11];
    const double P[9] = {0.0, -p2, p1, p2, 0.0, -p0, -p1, p0, 0.0};
    memset(Ad, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const double r = T[i * 4 + j];
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += P[i * 3 + k] * T[k * 4 + j];
            Ad[i * 6 + j] = r;
            Ad[(i + 3) * 6 + (j + 3)] = r;
            Ad[(i + 3) * 6 + j] = s;
        }
}

static void screw_col(const orc_robot *rb, int j, double *S6) {
    for (int r = 0; r < 6; ++r) S6[r] = rb->S[r * rb->n + j];
}

/* ------------------------------------------------------------------ kinematics */

/* kinematics/fk.py:61-70: T = prod_i exp([S_i] theta_i) @ M over the first k joints. */
static void fk_prefix(const orc_robot *rb, const double *theta, int k, double *T) {
    double E[16], S6[6];
    mat4_eye(T);
    for (int i = 0; i < k; ++i) {
        screw_col(rb, i, S6);
        orc_transform_from_twist(S6, theta[i], E);
        mat4_mul(T, E, T);
    }
    mat4_mul(T, rb->M, T);
}

void orc_fk_space(const orc_robot *rb, const double *theta, double *T) {
    fk_prefix(rb, theta, rb->n, T);
}

/* kinematics/jacobian.py:62-73: J[:, i] = Ad(T_{i-1}) S_i ; T *= exp([S_i] theta_i).
 * J is (6, n) row-major. */
void orc_jacobian_space(const orc_robot *rb, const double *theta, double *J) {
    const int n = rb->n;
    double T[16], E[16], Ad[36], S6[6];
    mat4_eye(T);
    for (int i = 0; i < n; ++i) {
        screw_col(rb, i, S6);
        orc_adjoint(T, Ad);
        for (int r = 0; r < 6; ++r) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += Ad[r * 6 + k] * S6[k];
            J[r * n + i] = s;
        }
        orc_transform_from_twist(S6, theta[i], E);
        mat4_mul(T, E, T);
    }
}

/* ------------------------------------------------------------------ literal dynamics */

/* Per-link CoM body Jacobian, shared by mass_matrix and gravity_forces
 * (dynamics/mass_matrix.py:66-91, dynamics/forces.py:106-121):
 *   T_k_com = FK(theta[:k+1]) @ inv(FK(0_{k+1})) @ Mlist_per_link[k]
 *   J_k     = Ad(inv(T_k_com)) @ J_s[:, :k+1]   (columns > k are zero). */
static void link_com_jacobian(const orc_robot *rb, const double *theta, const double *Js, int k,
                              double *Tkcom, double *Jk /* (6, n) */) {
    const int n = rb->n;
    double zeros[ORC_MAX_DOF] = {0};
    double Tk[16], Tk0[16], Tk0i[16], L2C[16], Ti[16], Ad[36];
    fk_prefix(rb, theta, k + 1, Tk);
    fk_prefix(rb, zeros, k + 1, Tk0);
    se3_inv(Tk0, Tk0i);
    mat4_mul(Tk0i, rb->Mcom + 16 * k, L2C);
    mat4_mul(Tk, L2C, Tkcom);
    se3_inv(Tkcom, Ti);
    orc_adjoint(Ti, Ad);
    memset(Jk, 0, 6 * n * sizeof(double));
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c <= k; ++c) {
            double s = 0.0;
            for (int q = 0; q < 6; ++q) s += Ad[r * 6 + q] * Js[q * n + c];
            Jk[r * n + c] = s;
        }
}

/* dynamics/mass_matrix.py:16-99: M = sym(sum_k J_k^T G_k J_k). Mout is (n, n). */
void orc_mass_matrix(const orc_robot *rb, const double *theta, double *Mout) {
    const int n = rb->n;
    double Js[6 * ORC_MAX_DOF], Jk[6 * ORC_MAX_DOF], GJ[6 * ORC_MAX_DOF], Tkc[16];
    double Macc[ORC_MAX_DOF * ORC_MAX_DOF];
    memset(Macc, 0, sizeof Macc);
    orc_jacobian_space(rb, theta, Js);
    for (int k = 0; k < n; ++k) {
        const double *G = rb->G + 36 * k;
        link_com_jacobian(rb, theta, Js, k, Tkc, Jk);
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < n; ++c) {
                double s = 0.0;
                for (int q = 0; q < 6; ++q) s += G[r * 6 + q] * Jk[q * n + c];
                GJ[r * n + c] = s;
            }
        for (int a = 0; a < n; ++a)
            for (int b = 0; b < n; ++b) {
                double s = 0.0;
                for (int q = 0; q < 6; ++q) s += Jk[q * n + a] * GJ[q * n + b];
                Macc[a * n + b] += s;
            }
    }
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) Mout[a * n + b] = 0.5 * (Macc[a * n + b] + Macc[b * n + a]);
}

/* dynamics/forces.py:61-133: g(theta) = sum_k J_k^T [0; G_k[3,3] R_k^T (-g)]. */
void orc_gravity_forces(const orc_robot *rb, const double *theta, const double *g3, double *out) {
    const int n = rb->n;
    double Js[6 * ORC_MAX_DOF], Jk[6 * ORC_MAX_DOF], Tkc[16];
    orc_jacobian_space(rb, theta, Js);
    for (int i = 0; i < n; ++i) out[i] = 0.0;
    for (int k = 0; k < n; ++k) {
        link_com_jacobian(rb, theta, Js, k, Tkc, Jk);
        const double m = rb->G[36 * k + 3 * 6 + 3];
        double F[6] = {0, 0, 0, 0, 0, 0};
        for (int r = 0; r < 3; ++r) {
            double s = 0.0;
            for (int q = 0; q < 3; ++q) s += Tkc[q * 4 + r] * (-g3[q]);
            F[3 + r] = m * s;
        }
        for (int c = 0; c < n; ++c) {
            double s = 0.0;
            for (int q = 0; q < 6; ++q) s += Jk[q * n + c] * F[q];
            out[c] += s;
        }
    }
}

/* dynamics/cache.py:23-56: dM[i, j, k] = (M(theta + eps e_k) - M(theta - eps e_k))_ij / (2 eps). */
static void mass_matrix_derivatives(const orc_robot *rb, const double *theta, double eps, double *dM) {
    const int n = rb->n;
    double tp[ORC_MAX_DOF], tm[ORC_MAX_DOF];
    double Mp[ORC_MAX_DOF * ORC_MAX_DOF], Mm[ORC_MAX_DOF * ORC_MAX_DOF];
    for (int k = 0; k < n; ++k) {
        for (int i = 0; i < n; ++i) {
            const double e = (i == k) ? 1.0 : 0.0;
            tp[i] = theta[i] + eps * e;
            tm[i] = theta[i] - eps * e;
        }
        orc_mass_matrix(rb, tp, Mp);
        orc_mass_matrix(rb, tm, Mm);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
                dM[(i * n + j) * n + k] = (Mp[i * n + j] - Mm[i * n + j]) / (2.0 * eps);
    }
}

/* dynamics/forces.py:26-59: c_i = dtheta^T Gamma_i dtheta,
 * Gamma_i[j,k] = 0.5 (dM[i,j,k] + dM[i,k,j] - dM[j,k,i]). */
void orc_velocity_quadratic_forces(const orc_robot *rb, const double *theta, const double *dtheta,
                                   double *c) {
    const int n = rb->n;
    double dM[ORC_MAX_DOF * ORC_MAX_DOF * ORC_MAX_DOF];
    mass_matrix_derivatives(rb, theta, 1e-6, dM);
    for (int i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int j = 0; j < n; ++j) {
            double row = 0.0;
            for (int k = 0; k < n; ++k) {
                const double gam =
                    0.5 * (dM[(i * n + j) * n + k] + dM[(i * n + k) * n + j] - dM[(j * n + k) * n + i]);
                row += gam * dtheta[k];
            }
            acc += dtheta[j] * row;
        }
        c[i] = acc;
    }
}

/* dynamics/id_fd.py:16-48: tau = M ddtheta + c + g + J_s^T Ftip. */
void orc_inverse_dynamics(const orc_robot *rb, const double *theta, const double *dtheta,
                          const double *ddtheta, const double *g3, const double *Ftip, double *tau) {
    const int n = rb->n;
    double Mm[ORC_MAX_DOF * ORC_MAX_DOF], c[ORC_MAX_DOF], gf[ORC_MAX_DOF], Js[6 * ORC_MAX_DOF];
    orc_mass_matrix(rb, theta, Mm);
    orc_velocity_quadratic_forces(rb, theta, dtheta, c);
    orc_gravity_forces(rb, theta, g3, gf);
    orc_jacobian_space(rb, theta, Js);
    for (int i = 0; i < n; ++i) {
        double md = 0.0, jf = 0.0;
        for (int j = 0; j < n; ++j) md += Mm[i * n + j] * ddtheta[j];
        for (int q = 0; q < 6; ++q) jf += Js[q * n + i] * Ftip[q];
        tau[i] = md + c[i] + gf[i] + jf;
    }
}

/* LU with partial pivoting (what np.linalg.solve / LAPACK gesv does). Returns 0 on success. */
static int lu_solve(int n, double *A, double *b) {
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i)
            if (fabs(A[i * n + k]) > best) { best = fabs(A[i * n + k]); piv = i; }
        if (best == 0.0) return -1;
        if (piv != k) {
            for (int j = 0; j < n; ++j) { double t = A[k * n + j]; A[k * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
            double t = b[k]; b[k] = b[piv]; b[piv] = t;
        }
        for (int i = k + 1; i < n; ++i) {
            const double l = A[i * n + k] / A[k * n + k];
            A[i * n + k] = l;
            for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
            b[i] -= l * b[k];
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < n; ++j) s -= A[i * n + j] * b[j];
        b[i] = s / A[i * n + i];
    }
    return 0;
}

/* dynamics/id_fd.py:50-83: ddtheta = solve(M, tau - c - g - J_s^T Ftip). */
int orc_forward_dynamics(const orc_robot *rb, const double *theta, const double *dtheta,
                         const double *tau, const double *g3, const double *Ftip, double *ddtheta) {
    const int n = rb->n;
    double Mm[ORC_MAX_DOF * ORC_MAX_DOF], c[ORC_MAX_DOF], gf[ORC_MAX_DOF], Js[6 * ORC_MAX_DOF];
    orc_mass_matrix(rb, theta, Mm);
    orc_velocity_quadratic_forces(rb, theta, dtheta, c);
    orc_gravity_forces(rb, theta, g3, gf);
    orc_jacobian_space(rb, theta, Js);
    for (int i = 0; i < n; ++i) {
        double jf = 0.0;
        for (int q = 0; q < 6; ++q) jf += Js[q * n + i] * Ftip[q];
        ddtheta[i] = tau[i] - c[i] - gf[i] - jf;
    }
    return lu_solve(n, Mm, ddtheta);
}

/* ------------------------------------------------------------------ analytic recursion
 * body-frame Newton-Euler in the link-CoM frames.
 *   A_i        = Ad(Mcom_i^-1) S_i
 *   T_{i,i-1}  = exp(-[A_i] theta_i) Mcom_i^-1 Mcom_{i-1}
 *   V_i  = Ad(T_{i,i-1}) V_{i-1} + A_i dth_i
 *   Vd_i = Ad(T_{i,i-1}) Vd_{i-1} + ad(V_i) A_i dth_i + A_i ddth_i
 *   F_i  = G_i Vd_i - ad(V_i)^T G_i V_i + [0; m_i R_i^T(-g)] + Ad(T_{i+1,i})^T F_{i+1}
 *   tau_i = F_i . A_i  (+ J_s^T Ftip)
 * The gravity wrench is added explicitly exactly as dynamics/forces.py:125-131 does,
 * so it is valid for any G_i; G_i is symmetrised like mass_matrix.py:96 implies.   */

static void mat6_vec(const double *A, const double *x, double *y) {
    for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += A[r * 6 + k] * x[k];
        y[r] = s;
    }
}
static void mat6T_vec(const double *A, const double *x, double *y) {
    for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += A[k * 6 + r] * x[k];
        y[r] = s;
    }
}
/* ad(V) = [[ [w], 0 ], [ [v], [w] ]] */
static void small_ad(const double *V, double *ad) {
    const double w0 = V[0], w1 = V[1], w2 = V[2], v0 = V[3], v1 = V[4], v2 = V[5];
    const double W[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
    const double U[9] = {0, -v2, v1, v2, 0, -v0, -v1, v0, 0};
    memset(ad, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            ad[i * 6 + j] = W[i * 3 + j];
            ad[(i + 3) * 6 + j + 3] = W[i * 3 + j];
            ad[(i + 3) * 6 + j] = U[i * 3 + j];
        }
}

void orc_rnea_analytic(const orc_robot *rb, const double *theta, const double *dtheta,
                       const double *ddtheta, const double *g3, const double *Ftip, double *tau) {
    const int n = rb->n;
    double A[ORC_MAX_DOF][6], AdT[ORC_MAX_DOF + 1][36], V[ORC_MAX_DOF + 1][6], Vd[ORC_MAX_DOF + 1][6];
    double Rw[ORC_MAX_DOF][9];
    double P[16], E[16], S6[6], Mi[16], Tmp[16], Ad[36];
    memset(V[0], 0, sizeof V[0]);
    memset(Vd[0], 0, sizeof Vd[0]);
    mat4_eye(P);
    for (int i = 0; i < n; ++i) {
        const double *Mc = rb->Mcom + 16 * i;
        screw_col(rb, i, S6);
        se3_inv(Mc, Mi);
        orc_adjoint(Mi, Ad);
        mat6_vec(Ad, S6, A[i]);
        /* T_{i,i-1} = exp(-[A_i] th) Mcom_i^-1 Mcom_{i-1} */
        double Tii[16], Ei[16];
        orc_transform_from_twist(A[i], -theta[i], Ei);
        if (i == 0) memcpy(Tmp, Mi, sizeof Tmp);
        else mat4_mul(Mi, rb->Mcom + 16 * (i - 1), Tmp);
        mat4_mul(Ei, Tmp, Tii);
        orc_adjoint(Tii, AdT[i]);
        /* world rotation of CoM frame i: P_i Mcom_i */
        orc_transform_from_twist(S6, theta[i], E);
        mat4_mul(P, E, P);
        mat4_mul(P, Mc, Tmp);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Rw[i][r * 3 + c] = Tmp[r * 4 + c];
        /* twists */
        double t6[6], adv[36], t6b[6];
        mat6_vec(AdT[i], V[i], t6);
        for (int r = 0; r < 6; ++r) V[i + 1][r] = t6[r] + A[i][r] * dtheta[i];
        mat6_vec(AdT[i], Vd[i], t6);
        small_ad(V[i + 1], adv);
        mat6_vec(adv, A[i], t6b);
        for (int r = 0; r < 6; ++r) Vd[i + 1][r] = t6[r] + t6b[r] * dtheta[i] + A[i][r] * ddtheta[i];
    }
    double F[6] = {0, 0, 0, 0, 0, 0};
    for (int i = n - 1; i >= 0; --i) {
        const double *G = rb->G + 36 * i;
        double Gs[36], GVd[6], GV[6], adv[36], adTGV[6], back[6];
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) Gs[r * 6 + c] = 0.5 * (G[r * 6 + c] + G[c * 6 + r]);
        mat6_vec(Gs, Vd[i + 1], GVd);
        mat6_vec(Gs, V[i + 1], GV);
        small_ad(V[i + 1], adv);
        mat6T_vec(adv, GV, adTGV);
        if (i == n - 1) memset(back, 0, sizeof back);
        else mat6T_vec(AdT[i + 1], F, back);
        const double m = G[3 * 6 + 3];
        for (int r = 0; r < 6; ++r) F[r] = GVd[r] - adTGV[r] + back[r];
        for (int r = 0; r < 3; ++r) {
            double s = 0.0;
            for (int q = 0; q < 3; ++q) s += Rw[i][q * 3 + r] * (-g3[q]);
            F[3 + r] += m * s;
        }
        double s = 0.0;
        for (int r = 0; r < 6; ++r) s += F[r] * A[i][r];
        tau[i] = s;
    }
    int has_tip = 0;
    for (int q = 0; q < 6; ++q) has_tip |= (Ftip != NULL && Ftip[q] != 0.0);
    if (has_tip) {
        double Js[6 * ORC_MAX_DOF];
        orc_jacobian_space(rb, theta, Js);
        for (int i = 0; i < n; ++i) {
            double jf = 0.0;
            for (int q = 0; q < 6; ++q) jf += Js[q * n + i] * Ftip[q];
            tau[i] += jf;
        }
    }
}

/* mass matrix column j = tau(theta, 0, e_j, g = 0, Ftip = 0); symmetrised. */
void orc_mass_matrix_analytic(const orc_robot *rb, const double *theta, double *Mout) {
    const int n = rb->n;
    const double z3[3] = {0, 0, 0}, z6[6] = {0, 0, 0, 0, 0, 0};
    double zero[ORC_MAX_DOF] = {0}, e[ORC_MAX_DOF], col[ORC_MAX_DOF];
    double Mm[ORC_MAX_DOF * ORC_MAX_DOF];
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < n; ++i) e[i] = (i == j) ? 1.0 : 0.0;
        orc_rnea_analytic(rb, theta, zero, e, z3, z6, col);
        for (int i = 0; i < n; ++i) Mm[i * n + j] = col[i];
    }
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) Mout[a * n + b] = 0.5 * (Mm[a * n + b] + Mm[b * n + a]);
}

int orc_forward_dynamics_analytic(const orc_robot *rb, const double *theta, const double *dtheta,
                                  const double *tau, const double *g3, const double *Ftip,
                                  double *ddtheta) {
    const int n = rb->n;
    double zero[ORC_MAX_DOF] = {0}, bias[ORC_MAX_DOF], Mm[ORC_MAX_DOF * ORC_MAX_DOF];
    orc_rnea_analytic(rb, theta, dtheta, zero, g3, Ftip, bias);
    orc_mass_matrix_analytic(rb, theta, Mm);
    for (int i = 0; i < n; ++i) ddtheta[i] = tau[i] - bias[i];
    return lu_solve(n, Mm, ddtheta);
}

/* The kernels' solve: LDL^T without pivoting (csrc/mpk_device.cuh ldlt_factor / ldlt_apply, same
 * operation order; the mass matrix is symmetric positive definite).  With it the oracle differs from a
 * rollout kernel only by roundings inside one operation (FMA contraction, the reciprocal of a pivot), so
 * what divergence is left after many chaotic steps is attributable to those, not to LU against LDL^T. */
static int ldlt_solve(int n, double *A, double *b) {
    double dinv[ORC_MAX_DOF];
    for (int j = 0; j < n; ++j) {
        double dj = A[j * n + j];
        for (int k = 0; k < j; ++k) dj -= A[j * n + k] * A[j * n + k] * A[k * n + k];
        if (dj == 0.0) return -1;
        A[j * n + j] = dj;
        dinv[j] = 1.0 / dj;
        for (int i = j + 1; i < n; ++i) {
            double l = A[i * n + j];
            for (int k = 0; k < j; ++k) l -= A[i * n + k] * A[j * n + k] * A[k * n + k];
            A[i * n + j] = l * dinv[j];
        }
    }
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < i; ++k) b[i] -= A[i * n + k] * b[k];
    for (int i = 0; i < n; ++i) b[i] *= dinv[i];
    for (int i = n - 1; i >= 0; --i)
        for (int k = i + 1; k < n; ++k) b[i] -= A[k * n + i] * b[k];
    return 0;
}

int orc_forward_dynamics_analytic_ldlt(const orc_robot *rb, const double *theta, const double *dtheta,
                                       const double *tau, const double *g3, const double *Ftip,
                                       double *ddtheta) {
    const int n = rb->n;
    double zero[ORC_MAX_DOF] = {0}, bias[ORC_MAX_DOF], Mm[ORC_MAX_DOF * ORC_MAX_DOF];
    orc_rnea_analytic(rb, theta, dtheta, zero, g3, Ftip, bias);
    orc_mass_matrix_analytic(rb, theta, Mm);
    for (int i = 0; i < n; ++i) ddtheta[i] = tau[i] - bias[i];
    return ldlt_solve(n, Mm, ddtheta);
}

/* ------------------------------------------------------------------ trajectory level */

static float clipf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* planning/trajectory.py:15-75 (_trajectory_cpu_fallback) + :311-313 (clip):
 * float64 time scaling, ONE rounding to float32 per element; tau = idx*(Tf/(N-1))/Tf.
 * method 3 = cubic, 5 = quintic, anything else = zero scaling (the planner's CPU
 * contract).  inputs_f32 = 1 restates joint_trajectory (planning/trajectory.py:147-153:
 * start/end cast to float32 first, so dtheta is a float32 subtraction);
 * inputs_f32 = 0 restates batch_joint_trajectory, which hands float64 rows to the same
 * Numba kernel without the cast (planning/trajectory.py:361-362, 474-476).
 * limits: (n, 2) float32 or NULL.  Outputs (N, n) float32 row-major. */
void orc_joint_trajectory(int n, const double *start, const double *end, int inputs_f32, double Tf,
                          int64_t N, int method, const float *limits, float *pos, float *vel,
                          float *acc) {
    for (int64_t idx = 0; idx < N; ++idx) {
        const double t = (double)idx * (Tf / (double)(N - 1));
        const double tau = t / Tf;
        double s, sd, sdd;
        if (method == 3) {
            s = 3.0 * tau * tau - 2.0 * tau * tau * tau;
            sd = 6.0 * tau * (1.0 - tau) / Tf;
            sdd = 6.0 / (Tf * Tf) * (1.0 - 2.0 * tau);
        } else if (method == 5) {
            const double t2 = tau * tau, t3 = t2 * tau, t4 = t2 * t2, t5 = t4 * tau;
            s = 10.0 * t3 - 15.0 * t4 + 6.0 * t5;
            sd = (30.0 * t2 - 60.0 * t3 + 30.0 * t4) / Tf;
            sdd = (60.0 * tau - 180.0 * t2 + 120.0 * t3) / (Tf * Tf);
        } else {
            s = sd = sdd = 0.0;
        }
        for (int j = 0; j < n; ++j) {
            double st, dth;
            if (inputs_f32) {
                const float s32 = (float)start[j], e32 = (float)end[j];
                st = (double)s32;
                dth = (double)(e32 - s32);
            } else {
                st = start[j];
                dth = end[j] - start[j];
            }
            float p = (float)(s * dth + st);
            if (limits) p = clipf(p, limits[2 * j], limits[2 * j + 1]);
            pos[idx * n + j] = p;
            vel[idx * n + j] = (float)(sd * dth);
            acc[idx * n + j] = (float)(sdd * dth);
        }
    }
}

/* planning/trajectory_dynamics.py:308-380 (_inverse_dynamics_cpu): per point ID in
 * float64, row cast to float32, then clip to float32 torque limits.
 * analytic = 0 -> literal finite-difference path, 1 -> analytic recursion (2, rollouts: ... with the
 * kernels' LDL^T solve instead of LU).
 * Ftip is one (6,) wrench for every point. */
void orc_inverse_dynamics_trajectory(const orc_robot *rb, int64_t P, const double *theta,
                                     const double *dtheta, const double *ddtheta, const double *g3,
                                     const double *Ftip, const float *tau_limits, int analytic,
                                     float *out) {
    const int n = rb->n;
    for (int64_t p = 0; p < P; ++p) {
        double tau[ORC_MAX_DOF];
        if (analytic)
            orc_rnea_analytic(rb, theta + p * n, dtheta + p * n, ddtheta + p * n, g3, Ftip, tau);
        else
            orc_inverse_dynamics(rb, theta + p * n, dtheta + p * n, ddtheta + p * n, g3, Ftip, tau);
        for (int j = 0; j < n; ++j) {
            float t = (float)tau[j];
            if (tau_limits) t = clipf(t, tau_limits[2 * j], tau_limits[2 * j + 1]);
            out[p * n + j] = t;
        }
    }
}

/* float64 per-point batch (the ManipulatorDynamics.inverse_dynamics API, batched). */
void orc_inverse_dynamics_batch(const orc_robot *rb, int64_t P, const double *theta,
                                const double *dtheta, const double *ddtheta, const double *g3,
                                const double *Ftip, int64_t ftip_stride, int analytic, double *out) {
    const int n = rb->n;
    const double z6[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t p = 0; p < P; ++p) {
        const double *F = Ftip ? Ftip + p * ftip_stride : z6;
        if (analytic)
            orc_rnea_analytic(rb, theta + p * n, dtheta + p * n, ddtheta + p * n, g3, F, out + p * n);
        else
            orc_inverse_dynamics(rb, theta + p * n, dtheta + p * n, ddtheta + p * n, g3, F, out + p * n);
    }
}

void orc_fk_jacobian_batch(const orc_robot *rb, int64_t P, const double *theta, double *T,
                           double *J) {
    const int n = rb->n;
    for (int64_t p = 0; p < P; ++p) {
        if (T) orc_fk_space(rb, theta + p * n, T + p * 16);
        if (J) orc_jacobian_space(rb, theta + p * n, J + p * 6 * n);
    }
}

void orc_mass_matrix_batch(const orc_robot *rb, int64_t P, const double *theta, int analytic,
                           double *Mout) {
    const int n = rb->n;
    for (int64_t p = 0; p < P; ++p) {
        if (analytic) orc_mass_matrix_analytic(rb, theta + p * n, Mout + p * n * n);
        else orc_mass_matrix(rb, theta + p * n, Mout + p * n * n);
    }
}

/* planning/trajectory_dynamics.py:580-708 (_forward_dynamics_cpu), one trajectory:
 *   row 0 = initial state, ddtheta row 0 = 0; for i in 1..N-1, intRes sub-steps of
 *   ddth = FD(th, dth, taumat[i], g, Ftipmat[i]); dth += ddth*dts; th += dth*dts;
 *   th = clip(th, lo32, hi32); record float32 rows, acceleration = last sub-step's.
 * State stays float64 (float64 inputs).  limits (n,2) float32 (may hold +-inf).
 * Ftipmat may be NULL (zeros).  Returns 0, or -1 if a solve failed (the reference
 * would log and skip that sub-step; not reproduced -- never happens for PD M).    */
int orc_forward_dynamics_trajectory(const orc_robot *rb, const double *theta0,
                                    const double *dtheta0, int64_t N, const double *taumat,
                                    const double *g3, const double *Ftipmat, double dt, int intRes,
                                    const float *limits, int analytic, float *pos, float *vel,
                                    float *acc) {
    const int n = rb->n;
    const double z6[6] = {0, 0, 0, 0, 0, 0};
    double th[ORC_MAX_DOF], dth[ORC_MAX_DOF], dd[ORC_MAX_DOF], last[ORC_MAX_DOF];
    int rc = 0;
    for (int j = 0; j < n; ++j) {
        th[j] = theta0[j];
        dth[j] = dtheta0[j];
        pos[j] = (float)th[j];
        vel[j] = (float)dth[j];
        acc[j] = 0.0f;
    }
    const double dts = dt / (double)intRes;
    for (int64_t i = 1; i < N; ++i) {
        for (int j = 0; j < n; ++j) last[j] = 0.0;
        for (int r = 0; r < intRes; ++r) {
            const double *F = Ftipmat ? Ftipmat + 6 * i : z6;
            int e = analytic == 2 ? orc_forward_dynamics_analytic_ldlt(rb, th, dth, taumat + i * n, g3, F, dd)
                    : analytic ? orc_forward_dynamics_analytic(rb, th, dth, taumat + i * n, g3, F, dd)
                               : orc_forward_dynamics(rb, th, dth, taumat + i * n, g3, F, dd);
            if (e) { rc = -1; continue; }
            for (int j = 0; j < n; ++j) {
                dth[j] = dth[j] + dd[j] * dts;
                th[j] = th[j] + dth[j] * dts;
                if (limits) {
                    const double lo = (double)limits[2 * j], hi = (double)limits[2 * j + 1];
                    th[j] = th[j] < lo ? lo : (th[j] > hi ? hi : th[j]);
                }
                last[j] = dd[j];
            }
        }
        for (int j = 0; j < n; ++j) {
            pos[i * n + j] = (float)th[j];
            vel[i * n + j] = (float)dth[j];
            acc[i * n + j] = (float)last[j];
        }
    }
    return rc;
}

/* Batched rollouts (the B200 extension: B independent trajectories). Layouts:
 * theta0/dtheta0 (B, n), taumat (B, N, n), Ftipmat (B, N, 6) or NULL, out (B, N, n). */
int orc_forward_dynamics_rollout_batch(const orc_robot *rb, int64_t B, const double *theta0,
                                       const double *dtheta0, int64_t N, const double *taumat,
                                       const double *g3, const double *Ftipmat, double dt,
                                       int intRes, const float *limits, int analytic, float *pos,
                                       float *vel, float *acc) {
    const int n = rb->n;
    int rc = 0;
    for (int64_t b = 0; b < B; ++b) {
        rc |= orc_forward_dynamics_trajectory(
            rb, theta0 + b * n, dtheta0 + b * n, N, taumat + b * N * n, g3,
            Ftipmat ? Ftipmat + b * N * 6 : NULL, dt, intRes, limits, analytic, pos + b * N * n,
            vel + b * N * n, acc + b * N * n);
    }
    return rc;
}

int orc_max_dof(void) { return ORC_MAX_DOF; }
