// hostcheck.cu -- TEST INFRASTRUCTURE ONLY.
//
// Executes the kernels' own per-thread templates (manipulapy_b200/csrc/mpk_device.cuh) on
// the host CPU, one point at a time, so that the algebra of the CUDA path (joint-aligned
// frames, Newton-Euler recursion, CRBA, LDL^T, time scaling) can be checked against the
// oracle and the reference's golden vectors in the GPU-less build container.  Nothing in
// manipulapy_b200/ links or loads this library; it is not a CPU fallback.
#include "../../manipulapy_b200/csrc/mpk_common.cuh"

using namespace mpk;

// flavour of the kernels this robot is routed to (csrc/dyn_kernels.cuh): 0 rigid + all
// revolute, 1 rigid, 2 general inertias
static int flavour(const mpk_robot *rb) { return (!rb->rigid || !rb->first_revolute) ? 2 : (rb->plain ? 0 : 1); }

template <int N, bool GEN, bool REV, unsigned GEO = 0>
static void rnea_nf(const mpk_robot *rb, int64_t P, const double *th, const double *dth,
                    const double *ddth, const double *g, const double *ftip, double *tau,
                    int use_smem_store) {
    const RobotPack<double, N> pk = narrow<N>(rb);
    double g0[3];
    base_gravity(pk, g, g0);
    for (int64_t p = 0; p < P; ++p) {
        double a[N], b[N], c[N], t[N];
        for (int j = 0; j < N; ++j) {
            a[j] = th[p * N + j];
            b[j] = dth ? dth[p * N + j] : 0.0;
            c[j] = ddth ? ddth[p * N + j] : 0.0;
        }
        if (!GEN && !dth && !ddth && !ftip) {
            // gravity forces: the at-rest form of the recursion, as launch_rnea routes them
            using Store = SmemStore<double, N, 1, rnea_fast0(GEN, REV, N)>;
            double buf[Store::kValues + 1];
            Store st{buf};
            ArrayInAtRest<double, N> in{a};
            rnea<double, N, GEN, REV, GEO>(pk, in, g0, ftip, t, st);
        } else if (use_smem_store) {
            // the shared-memory state store of the kernels, exercised with a one-thread "block"
            using Store = SmemStore<double, N, 1, rnea_fast0(GEN, REV, N)>;
            double buf[Store::kValues + 1];
            Store st{buf};
            ArrayIn<double, N> in{a, b, c};
            rnea<double, N, GEN, REV, GEO>(pk, in, g0, ftip, t, st);
        } else {
            JointCS<double, N> q;
            rnea<double, N, GEN, REV, GEO>(pk, a, b, c, g0, ftip, t, q);
        }
        for (int j = 0; j < N; ++j) tau[p * N + j] = t[j];
    }
}

template <int N>
static void rnea_n(const mpk_robot *rb, int64_t P, const double *th, const double *dth,
                   const double *ddth, const double *g, const double *ftip, double *tau,
                   int use_smem_store) {
    // the kernels compiled for this robot's link-geometry signature, exactly as the launchers pick them
#define X(n_, g_)                                                                                    \
    if constexpr (N == n_)                                                                           \
        if (flavour(rb) == 0 && rb->geo == g_)                                                       \
            return rnea_nf<N, false, true, g_>(rb, P, th, dth, ddth, g, ftip, tau, use_smem_store);
    MPK_GEO_LIST(X)
#undef X
    switch (flavour(rb)) {
        case 0: rnea_nf<N, false, true>(rb, P, th, dth, ddth, g, ftip, tau, use_smem_store); break;
        case 1: rnea_nf<N, false, false>(rb, P, th, dth, ddth, g, ftip, tau, use_smem_store); break;
        default: rnea_nf<N, true, false>(rb, P, th, dth, ddth, g, ftip, tau, use_smem_store); break;
    }
}

template <int N, bool GEN, bool REV, unsigned GEO = 0>
static void mass_nf(const mpk_robot *rb, int64_t P, const double *th, double *Mo) {
    const RobotPack<double, N> pk = narrow<N>(rb);
    for (int64_t p = 0; p < P; ++p) {
        double a[N], Mm[N][N];
        for (int j = 0; j < N; ++j) a[j] = th[p * N + j];
        JointCS<double, N> q;
        joint_cs<double, N, REV>(pk, a, q);
        mass_matrix<double, N, GEN, REV, GEO>(pk, a, q, Mm);
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) Mo[(p * N + i) * N + j] = Mm[i][j];
    }
}

template <int N>
static void mass_n(const mpk_robot *rb, int64_t P, const double *th, double *Mo) {
#define X(n_, g_)             \
    if constexpr (N == n_)    \
        if (flavour(rb) == 0 && rb->geo == g_) return mass_nf<N, false, true, g_>(rb, P, th, Mo);
    MPK_GEO_LIST(X)
#undef X
    switch (flavour(rb)) {
        case 0: mass_nf<N, false, true>(rb, P, th, Mo); break;
        case 1: mass_nf<N, false, false>(rb, P, th, Mo); break;
        default: mass_nf<N, true, false>(rb, P, th, Mo); break;
    }
}

template <int N>
static void fk_n(const mpk_robot *rb, int64_t P, const double *th, double *T, double *J, bool body = false) {
    const RobotPack<double, N> pk = narrow<N>(rb);
    for (int64_t p = 0; p < P; ++p) {
        double a[N];
        for (int j = 0; j < N; ++j) a[j] = th[p * N + j];
        JointCS<double, N> q;
        joint_cs(pk, a, q);
        fk_jacobian<double, N>(pk, q, T ? T + p * 16 : nullptr, J ? J + p * 6 * N : nullptr, body);
    }
}

template <int N, bool GEN, bool REV, unsigned GEO = 0>
static void fd_nf(const mpk_robot *rb, int64_t P, const double *th, const double *dth,
                  const double *tau, const double *g, const double *ftip_rows, double *dd) {
    const RobotPack<double, N> pk = narrow<N>(rb);
    double g0[3];
    base_gravity(pk, g, g0);
    for (int64_t p = 0; p < P; ++p) {
        double a[N], b[N], c[N], o[N];
        for (int j = 0; j < N; ++j) {
            a[j] = th[p * N + j];
            b[j] = dth[p * N + j];
            c[j] = tau[p * N + j];
        }
        const double *ft = ftip_rows ? ftip_rows + 6 * p : nullptr;
        forward_dynamics<double, N, GEN, REV, 0, GEO>(pk, a, b, c, g0, ft, o);
        for (int j = 0; j < N; ++j) dd[p * N + j] = o[j];
    }
}

template <int N>
static void fd_n(const mpk_robot *rb, int64_t P, const double *th, const double *dth,
                 const double *tau, const double *g, const double *ftip_rows, double *dd) {
#define X(n_, g_)                                    \
    if constexpr (N == n_)                           \
        if (flavour(rb) == 0 && rb->geo == g_)       \
            return fd_nf<N, false, true, g_>(rb, P, th, dth, tau, g, ftip_rows, dd);
    MPK_GEO_LIST(X)
#undef X
    switch (flavour(rb)) {
        case 0: fd_nf<N, false, true>(rb, P, th, dth, tau, g, ftip_rows, dd); break;
        case 1: fd_nf<N, false, false>(rb, P, th, dth, tau, g, ftip_rows, dd); break;
        default: fd_nf<N, true, false>(rb, P, th, dth, tau, g, ftip_rows, dd); break;
    }
}

#define HC_DISPATCH(n, ...)                                      \
    switch (n) {                                                 \
        case 1: { constexpr int N_ = 1; __VA_ARGS__; } break;    \
        case 2: { constexpr int N_ = 2; __VA_ARGS__; } break;    \
        case 3: { constexpr int N_ = 3; __VA_ARGS__; } break;    \
        case 4: { constexpr int N_ = 4; __VA_ARGS__; } break;    \
        case 5: { constexpr int N_ = 5; __VA_ARGS__; } break;    \
        case 6: { constexpr int N_ = 6; __VA_ARGS__; } break;    \
        case 7: { constexpr int N_ = 7; __VA_ARGS__; } break;    \
        case 8: { constexpr int N_ = 8; __VA_ARGS__; } break;    \
        default: return -2;                                      \
    }

// Route a robot through the general kernels (signature 0) or back through the ones of its own
// link-geometry signature: the tests compare the two.  Returns the previous signature.
extern "C" unsigned hc_set_geo(mpk_robot *rb, unsigned geo) {
    const unsigned old = rb->geo;
    rb->geo = geo;
    return old;
}

extern "C" int hc_rnea(const mpk_robot *rb, int64_t P, const double *th, const double *dth,
                       const double *ddth, const double *g, const double *ftip, double *tau) {
    HC_DISPATCH(rb->n, rnea_n<N_>(rb, P, th, dth, ddth, g, ftip, tau, 0));
    return 0;
}
extern "C" int hc_rnea_smem(const mpk_robot *rb, int64_t P, const double *th, const double *dth,
                            const double *ddth, const double *g, const double *ftip, double *tau) {
    HC_DISPATCH(rb->n, rnea_n<N_>(rb, P, th, dth, ddth, g, ftip, tau, 1));
    return 0;
}
extern "C" int hc_mass(const mpk_robot *rb, int64_t P, const double *th, double *Mo) {
    HC_DISPATCH(rb->n, mass_n<N_>(rb, P, th, Mo));
    return 0;
}
extern "C" int hc_fk(const mpk_robot *rb, int64_t P, const double *th, double *T, double *J) {
    HC_DISPATCH(rb->n, fk_n<N_>(rb, P, th, T, J));
    return 0;
}
extern "C" int hc_fk_body(const mpk_robot *rb, int64_t P, const double *th, double *T, double *J) {
    HC_DISPATCH(rb->n, fk_n<N_>(rb, P, th, T, J, true));
    return 0;
}
extern "C" int hc_fd(const mpk_robot *rb, int64_t P, const double *th, const double *dth,
                     const double *tau, const double *g, const double *ftip_rows, double *dd) {
    HC_DISPATCH(rb->n, fd_n<N_>(rb, P, th, dth, tau, g, ftip_rows, dd));
    return 0;
}
// float32 arithmetic (the f32 kernel variants): same templates with T = float
template <int N, bool GEN, bool REV>
static void rnea32_nf(const mpk_robot *rb, int64_t P, const double *th, const double *dth,
                      const double *ddth, const double *g, const double *ftip, double *tau) {
    const RobotPack<float, N> pk = narrow<N, float>(rb);
    const RobotPack<double, N> pk64 = narrow<N>(rb);
    double g064[3];
    base_gravity(pk64, g, g064);  // (the launchers compute it in float64 on the host)
    float g0[3] = {(float)g064[0], (float)g064[1], (float)g064[2]};
    float ft[6];
    if (ftip)
        for (int k = 0; k < 6; ++k) ft[k] = (float)ftip[k];
    for (int64_t p = 0; p < P; ++p) {
        float a[N], b[N], c[N], t[N];
        for (int j = 0; j < N; ++j) {
            a[j] = (float)th[p * N + j];
            b[j] = dth ? (float)dth[p * N + j] : 0.f;
            c[j] = ddth ? (float)ddth[p * N + j] : 0.f;
        }
        using Store = SmemStore<float, N, 1, rnea_fast0(GEN, REV, N)>;
        float buf[Store::kValues + 1];
        Store st{buf};
        ArrayIn<float, N> in{a, b, c};
        rnea<float, N, GEN, REV>(pk, in, g0, ftip ? ft : nullptr, t, st);
        for (int j = 0; j < N; ++j) tau[p * N + j] = (double)t[j];
    }
}
template <int N>
static void fk32_n(const mpk_robot *rb, int64_t P, const double *th, double *T, double *J) {
    const RobotPack<float, N> pk = narrow<N, float>(rb);
    for (int64_t p = 0; p < P; ++p) {
        float a[N], To[16], Jo[6 * N];
        for (int j = 0; j < N; ++j) a[j] = (float)th[p * N + j];
        JointCS<float, N> q;
        joint_cs(pk, a, q);
        fk_jacobian<float, N>(pk, q, To, Jo);
        for (int k = 0; k < 16; ++k) T[p * 16 + k] = To[k];
        for (int k = 0; k < 6 * N; ++k) J[p * 6 * N + k] = Jo[k];
    }
}
extern "C" int hc_rnea_f32(const mpk_robot *rb, int64_t P, const double *th, const double *dth,
                           const double *ddth, const double *g, const double *ftip, double *tau) {
    HC_DISPATCH(rb->n, {
        switch (flavour(rb)) {
            case 0: rnea32_nf<N_, false, true>(rb, P, th, dth, ddth, g, ftip, tau); break;
            case 1: rnea32_nf<N_, false, false>(rb, P, th, dth, ddth, g, ftip, tau); break;
            default: rnea32_nf<N_, true, false>(rb, P, th, dth, ddth, g, ftip, tau); break;
        }
    });
    return 0;
}
extern "C" int hc_fk_f32(const mpk_robot *rb, int64_t P, const double *th, double *T, double *J) {
    HC_DISPATCH(rb->n, fk32_n<N_>(rb, P, th, T, J));
    return 0;
}
template <int N>
static void ik_n(const mpk_robot *rb, int64_t P, const double *Td, const double *th0,
                 const IkParams<double, MPK_MAX_DOF> &prm, unsigned long long seed, double *theta, int *iters,
                 unsigned char *ok, const double *noise, int noise_rows, int *restarts) {
    const RobotPack<double, N> pk = narrow<N>(rb);
    for (int64_t p = 0; p < P; ++p) {
        double th[N], J[6 * N + 1];
        for (int j = 0; j < N; ++j) th[j] = th0[p * N + j];
        int it = 0;
        int rs = 0;
        ok[p] = ik_dls<double, N>(pk, Td + 16 * p, th, prm, seed, (unsigned long long)p, J, it,
                                  noise ? noise + p * noise_rows * N : nullptr, noise_rows, &rs) ? 1 : 0;
        if (restarts) restarts[p] = rs;
        iters[p] = it;
        for (int j = 0; j < N; ++j) theta[p * N + j] = th[j];
    }
}
extern "C" int hc_ik(const mpk_robot *rb, int64_t P, const double *Td, const double *th0, double eomg, double ev,
                     int max_iterations, double damping, double step_cap, double w_rot, double w_pos,
                     const double *limits, unsigned long long seed, double *theta, int *iters,
                     unsigned char *ok, int flags, const double *noise, int noise_rows, int *restarts) {
    const IkParams<double, MPK_MAX_DOF> prm =
        make_ik_params(rb->n, eomg, ev, max_iterations, damping, step_cap, w_rot, w_pos, limits, flags);
    HC_DISPATCH(rb->n, ik_n<N_>(rb, P, Td, th0, prm, seed, theta, iters, ok, noise, noise_rows, restarts));
    return 0;
}
extern "C" int hc_cartesian(int64_t N, const double *Xs, const double *Xe, double Tf, int method, float *pos,
                            float *vel, float *acc, float *orient) {
    for (int64_t t = 0; t < N; ++t) {
        float p[3], v[3], a[3], R[9];
        cartesian_point(Xs, Xe, t, N, Tf, method, p, v, a, R);
        for (int k = 0; k < 3; ++k) {
            pos[3 * t + k] = p[k];
            vel[3 * t + k] = v[k];
            acc[3 * t + k] = a[k];
        }
        for (int k = 0; k < 9; ++k) orient[9 * t + k] = R[k];
    }
    return 0;
}
extern "C" int hc_sincos(int64_t P, const double *x, double *sn, double *cs) {
    double tab[17];
    fill_trig_table(tab);
    for (int64_t p = 0; p < P; ++p) sincos_pack(tab, x[p], sn + p, cs + p);
    return 0;
}
extern "C" int hc_traj(int n, int64_t N, const double *start, const double *end, int inputs_f32,
                       double Tf, int method, const float *limits, float *pos, float *vel, float *acc) {
    for (int64_t t = 0; t < N; ++t) {
        const TimeScale ts = time_scaling(t, N, Tf, method);
        for (int j = 0; j < n; ++j) {
            double st, dth;
            if (inputs_f32) {
                const float s32 = (float)start[j], e32 = (float)end[j];
                st = (double)s32;
                dth = (double)rn_fsub(e32, s32);
            } else {
                st = start[j];
                dth = rn_sub(end[j], start[j]);
            }
            traj_point(ts, st, dth, limits ? limits[2 * j] : 0.f, limits ? limits[2 * j + 1] : 0.f,
                       limits != nullptr, pos[t * n + j], vel[t * n + j], acc[t * n + j]);
        }
    }
    return 0;
}
