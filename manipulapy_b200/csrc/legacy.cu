// legacy.cu -- the reference's LEGACY dynamics path, batched: what a
// ManipulatorDynamics built by hand WITHOUT Mlist_per_link computes (tests/test_dynamics.py:59-67 builds
// one; the control loops of control/computed_torque.py:80-83 are single-sample consumers of it).
//
//   T_i = e^{[S_1] th_1} ... e^{[S_i] th_i} M        (kinematics/fk.py:61-70 on theta[:i + 1]: the END-EFFECTOR
//                                                    home pose for every link)
//   mass matrix   row i = J[:, i]^T (Ad(T_i)^T G_i Ad(T_i)) J,  then 0.5 (M + M^T)   (dynamics/mass_matrix.py:101-132)
//   gravity       g_i = (R_i^T g) . colsum(G_i[:3, :3])                              (dynamics/forces.py:135-154)
//   Coriolis      c_i = dth^T Gamma_i dth, Gamma from central differences of the mass matrix with
//                 eps = 1e-6 (dynamics/forces.py:26-59, dynamics/cache.py:23-56): 2 n mass matrices per point
//   inverse dynamics  M ddth + c + g + J^T Ftip   (dynamics/id_fd.py:16-48)
//   forward dynamics  solve(M, tau - c - g - J^T Ftip), LU with partial pivoting like np.linalg.solve (:50-83)
//
// The reference documents this path as incorrect physics ("DO NOT USE"); it is here, behind an explicit
// opt-in of the Python mirror, so that callers who still construct the object that way get the reference's
// numbers.  One thread owns one point; nothing here is tuned (the path has no performance claim).
#include <cstring>

#include "mpk_common.cuh"

namespace mpk {

struct LegacyPack {
    double S[6][MPK_MAX_DOF];
    double M[16];
    double G[MPK_MAX_DOF][36];
};

struct LegacyArgs {
    int64_t P;
    int mode;
    const double *th, *dth, *third;  // third: ddtheta (mode 3) or tau (mode 4)
    double g[3];
    double ftip[6];
    const double *ftip_rows;
    double *out;
};

struct Se3 {
    double R[9], p[3];
};

__device__ __forceinline__ Se3 se3_mul(const Se3 &a, const Se3 &b) {
    Se3 o;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) o.R[3 * r + c] = a.R[3 * r] * b.R[c] + a.R[3 * r + 1] * b.R[3 + c] + a.R[3 * r + 2] * b.R[6 + c];
        o.p[r] = a.p[r] + a.R[3 * r] * b.p[0] + a.R[3 * r + 1] * b.p[1] + a.R[3 * r + 2] * b.p[2];
    }
    return o;
}

// utils/se3.py:33-42 transform_from_twist: R = 1 + sin W + (1 - cos) W^2,  p = (1 th + (1 - cos) W + (th - sin) W^2) v
__device__ __forceinline__ Se3 exp_twist(const double (&w)[3], const double (&v)[3], double th) {
    double s, c;
    sincos(th, &s, &c);
    const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double W2[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) W2[3 * r + k] = W[3 * r] * W[k] + W[3 * r + 1] * W[3 + k] + W[3 * r + 2] * W[6 + k];
    Se3 o;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double pr = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double id = r == k ? 1.0 : 0.0;
            o.R[3 * r + k] = id + s * W[3 * r + k] + (1.0 - c) * W2[3 * r + k];
            pr += (id * th + (1.0 - c) * W[3 * r + k] + (th - s) * W2[3 * r + k]) * v[k];
        }
        o.p[r] = pr;
    }
    return o;
}

// y = Ad(T) x  (twists [w; v]: [R 0; [p] R  R])
__device__ __forceinline__ void adj_mul(const Se3 &T, const double (&x)[6], double (&y)[6]) {
    double rw[3], rv[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        rw[r] = T.R[3 * r] * x[0] + T.R[3 * r + 1] * x[1] + T.R[3 * r + 2] * x[2];
        rv[r] = T.R[3 * r] * x[3] + T.R[3 * r + 1] * x[4] + T.R[3 * r + 2] * x[5];
    }
    y[0] = rw[0]; y[1] = rw[1]; y[2] = rw[2];
    y[3] = T.p[1] * rw[2] - T.p[2] * rw[1] + rv[0];
    y[4] = T.p[2] * rw[0] - T.p[0] * rw[2] + rv[1];
    y[5] = T.p[0] * rw[1] - T.p[1] * rw[0] + rv[2];
}
// y = Ad(T)^T x
__device__ __forceinline__ void adjT_mul(const Se3 &T, const double (&x)[6], double (&y)[6]) {
    // Ad^T = [R^T  -R^T [p]; 0  R^T]:  top = R^T (x_w - p x x_v),  bottom = R^T x_v
    const double u[3] = {x[0] - (T.p[1] * x[5] - T.p[2] * x[4]), x[1] - (T.p[2] * x[3] - T.p[0] * x[5]),
                         x[2] - (T.p[0] * x[4] - T.p[1] * x[3])};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        y[c] = T.R[c] * u[0] + T.R[3 + c] * u[1] + T.R[6 + c] * u[2];
        y[3 + c] = T.R[c] * x[3] + T.R[3 + c] * x[4] + T.R[6 + c] * x[5];
    }
}

template <int N>
struct LegacyState {
    Se3 T[N];        // T_i = prefix_i M
    double J[6][N];  // space Jacobian
};

template <int N>
__device__ void legacy_frames(const LegacyPack &pk, const double (&th)[N], LegacyState<N> &st) {
    Se3 P;
#pragma unroll
    for (int k = 0; k < 9; ++k) P.R[k] = (k % 4 == 0) ? 1.0 : 0.0;
    P.p[0] = P.p[1] = P.p[2] = 0.0;
    Se3 Mh;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Mh.R[3 * r + c] = pk.M[4 * r + c];
        Mh.p[r] = pk.M[4 * r + 3];
    }
    for (int i = 0; i < N; ++i) {
        double S[6], col[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) S[k] = pk.S[k][i];
        adj_mul(P, S, col);
#pragma unroll
        for (int k = 0; k < 6; ++k) st.J[k][i] = col[k];
        const double w[3] = {S[0], S[1], S[2]}, v[3] = {S[3], S[4], S[5]};
        Se3 E;
        if (w[0] == 0.0 && w[1] == 0.0 && w[2] == 0.0) {
            // (transform_from_twist of a prismatic screw: R = 1, p = v th -- the same formula evaluates to it)
#pragma unroll
            for (int k = 0; k < 9; ++k) E.R[k] = (k % 4 == 0) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) E.p[k] = v[k] * th[i];
        } else {
            E = exp_twist(w, v, th[i]);
        }
        P = se3_mul(P, E);
        st.T[i] = se3_mul(P, Mh);
    }
}

template <int N>
__device__ void legacy_mass(const LegacyPack &pk, const double (&th)[N], double (&Mm)[N][N]) {
    LegacyState<N> st;
    legacy_frames<N>(pk, th, st);
    for (int i = 0; i < N; ++i) {
        double Ji[6], u[6], w[6], v[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) Ji[k] = st.J[k][i];
        // row i = J_i^T (Ad^T G Ad) J = (Ad^T G^T Ad J_i)^T J
        adj_mul(st.T[i], Ji, u);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) s += pk.G[i][6 * k + r] * u[k];
            w[r] = s;
        }
        adjT_mul(st.T[i], w, v);
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) s += v[k] * st.J[k][j];
            Mm[i][j] = s;
        }
    }
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) {
            const double a = 0.5 * (Mm[i][j] + Mm[j][i]);
            Mm[i][j] = a;
            Mm[j][i] = a;
        }
}

template <int N>
__device__ void legacy_gravity(const LegacyPack &pk, const LegacyState<N> &st, const double (&g)[3], double (&out)[N]) {
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double rg = st.T[i].R[c] * g[0] + st.T[i].R[3 + c] * g[1] + st.T[i].R[6 + c] * g[2];  // (R^T g)_c
            const double col = pk.G[i][c] + pk.G[i][6 + c] + pk.G[i][12 + c];                             // column sum of G[:3, :3]
            s += rg * col;
        }
        out[i] = s;
    }
}

template <int N>
__device__ void legacy_coriolis(const LegacyPack &pk, const double (&th)[N], const double (&dth)[N], double (&c)[N]) {
    const double eps = 1e-6;
    double dM[N][N][N];  // dM[i][j][k] = d M_ij / d th_k
    for (int k = 0; k < N; ++k) {
        double tp[N], tm[N], Mp[N][N], Mn[N][N];
        for (int j = 0; j < N; ++j) {
            tp[j] = th[j] + (j == k ? eps : 0.0);
            tm[j] = th[j] - (j == k ? eps : 0.0);
        }
        legacy_mass<N>(pk, tp, Mp);
        legacy_mass<N>(pk, tm, Mn);
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) dM[i][j][k] = (Mp[i][j] - Mn[i][j]) / (2.0 * eps);
    }
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
        for (int j = 0; j < N; ++j) {
            double row = 0.0;
            for (int k = 0; k < N; ++k) row += 0.5 * (dM[i][j][k] + dM[i][k][j] - dM[j][k][i]) * dth[k];
            s += dth[j] * row;
        }
        c[i] = s;
    }
}

// LU with partial pivoting (np.linalg.solve); b <- A^-1 b
template <int N>
__device__ void lu_solve(double (&A)[N][N], double (&b)[N]) {
    for (int k = 0; k < N; ++k) {
        int piv = k;
        double best = fabs(A[k][k]);
        for (int i = k + 1; i < N; ++i)
            if (fabs(A[i][k]) > best) {
                best = fabs(A[i][k]);
                piv = i;
            }
        if (piv != k) {
            for (int j = 0; j < N; ++j) {
                const double t = A[k][j];
                A[k][j] = A[piv][j];
                A[piv][j] = t;
            }
            const double t = b[k];
            b[k] = b[piv];
            b[piv] = t;
        }
        for (int i = k + 1; i < N; ++i) {
            const double l = A[i][k] / A[k][k];
            for (int j = k + 1; j < N; ++j) A[i][j] -= l * A[k][j];
            b[i] -= l * b[k];
        }
    }
    for (int i = N - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < N; ++j) s -= A[i][j] * b[j];
        b[i] = s / A[i][i];
    }
}

template <int N>
__global__ void __launch_bounds__(64) legacy_dynamics_kernel(const __grid_constant__ LegacyPack pk, const LegacyArgs a) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.P) return;
    double th[N], dth[N], x3[N];
    for (int j = 0; j < N; ++j) {
        th[j] = a.th[p * N + j];
        dth[j] = a.dth ? a.dth[p * N + j] : 0.0;
        x3[j] = a.third ? a.third[p * N + j] : 0.0;
    }
    if (a.mode == 0) {
        double Mm[N][N];
        legacy_mass<N>(pk, th, Mm);
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) a.out[(p * N + i) * N + j] = Mm[i][j];
        return;
    }
    LegacyState<N> st;
    legacy_frames<N>(pk, th, st);
    double out[N];
    if (a.mode == 1) {
        legacy_gravity<N>(pk, st, a.g, out);
    } else if (a.mode == 2) {
        legacy_coriolis<N>(pk, th, dth, out);
    } else {
        double grav[N], cor[N], Mm[N][N], ft[6];
        legacy_gravity<N>(pk, st, a.g, grav);
        legacy_coriolis<N>(pk, th, dth, cor);
        legacy_mass<N>(pk, th, Mm);
        for (int k = 0; k < 6; ++k) ft[k] = a.ftip_rows ? a.ftip_rows[p * 6 + k] : a.ftip[k];
        for (int i = 0; i < N; ++i) {
            double jf = 0.0;
            for (int k = 0; k < 6; ++k) jf += st.J[k][i] * ft[k];
            if (a.mode == 3) {
                double md = 0.0;
                for (int j = 0; j < N; ++j) md += Mm[i][j] * x3[j];
                out[i] = md + cor[i] + grav[i] + jf;
            } else {
                out[i] = x3[i] - cor[i] - grav[i] - jf;
            }
        }
        if (a.mode == 4) lu_solve<N>(Mm, out);
    }
    for (int j = 0; j < N; ++j) a.out[p * N + j] = out[j];
}

}  // namespace mpk

using namespace mpk;

extern "C" int mpk_legacy_dynamics(int n, const double *S_list, const double *M, const double *Glist, int mode,
                                   int64_t P, const double *theta, const double *dtheta, const double *third,
                                   const double *g, const double *Ftip, const double *Ftip_rows, double *out,
                                   void *stream) {
    if (n < 1 || n > MPK_MAX_DOF) return fail(MPK_EUNSUPPORTED, "dof must be in 1..8");
    if (!S_list || !M || !Glist) return fail(MPK_EINVAL, "S_list, M and Glist are required");
    if (mode < 0 || mode > 4) return fail(MPK_EINVAL, "mode must be 0..4");
    if (P < 0) return fail(MPK_EINVAL, "negative size");
    if (P == 0) return MPK_OK;
    if (!theta || !out) return fail(MPK_EINVAL, "theta and out are required");
    if (mode >= 2 && !dtheta) return fail(MPK_EINVAL, "dtheta is required");
    if (mode >= 3 && !third) return fail(MPK_EINVAL, "ddtheta / tau is required");
    LegacyPack pk;
    std::memset(&pk, 0, sizeof pk);
    for (int k = 0; k < 6; ++k)
        for (int i = 0; i < n; ++i) pk.S[k][i] = S_list[k * n + i];
    for (int k = 0; k < 16; ++k) pk.M[k] = M[k];
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 36; ++k) pk.G[i][k] = Glist[36 * i + k];
    LegacyArgs a;
    a.P = P;
    a.mode = mode;
    a.th = theta;
    a.dth = dtheta;
    a.third = third;
    for (int k = 0; k < 3; ++k) a.g[k] = g ? g[k] : 0.0;
    for (int k = 0; k < 6; ++k) a.ftip[k] = Ftip ? Ftip[k] : 0.0;
    a.ftip_rows = Ftip_rows;
    a.out = out;
    const int64_t blocks = (P + 63) / 64;
    if (blocks > 0x7fffffffLL) return fail(MPK_EINVAL, "P exceeds the grid limit");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MPK_DISPATCH_DOF(n, (legacy_dynamics_kernel<N_><<<(unsigned)blocks, 64, 0, s>>>(pk, a)));
    return check_launch("legacy_dynamics");
}
