// dyn.cu -- C ABI of the dynamics launchers: argument checking, argument blocks, dispatch to
// the flavour translation units (dyn_kernels.cuh).
#include "dyn_kernels.cuh"

using namespace mpk;

namespace {

TipArgs make_tip(const mpk_robot *rb, const double *g, const double *Ftip, const double *Ftip_rows) {
    TipArgs t;
    const double gz[3] = {g ? g[0] : 0.0, g ? g[1] : 0.0, g ? g[2] : 0.0};
    base_gravity(rb->pack, gz, t.g0);
    t.has_ftip = 0;
    for (int k = 0; k < 6; ++k) {
        t.ftip[k] = Ftip ? Ftip[k] : 0.0;
        if (t.ftip[k] != 0.0) t.has_ftip = 1;
    }
    t.ftip_rows = Ftip_rows;
    return t;
}

int grid_for(int64_t P, unsigned &grid) {
    const int64_t blocks = (P + kDynThreads - 1) / kDynThreads;
    if (blocks > 0x7fffffffLL) return fail(MPK_EINVAL, "point count exceeds the grid limit");
    grid = (unsigned)blocks;
    return MPK_OK;
}

}  // namespace

#define MPK_REQUIRE_DYN(rb)                                                      \
    if (!(rb)) return fail(MPK_EINVAL, "robot is NULL");                         \
    if (!(rb)->has_dynamics)                                                     \
        return fail(MPK_EINVAL, "robot was created without Glist / Mlist_per_link")

static int inverse_dynamics_impl(const mpk_robot *rb, int64_t P, const void *theta,
                                 const void *dtheta, const void *ddtheta, int in_dtype,
                                 const double *g, const double *Ftip, const double *Ftip_rows,
                                 const float *tau_limits, void *tau, int out_dtype, int compute_f32,
                                 void *stream) {
    MPK_REQUIRE_DYN(rb);
    if (P < 0) return fail(MPK_EINVAL, "negative size");
    if (P == 0) return MPK_OK;
    if (!theta || !tau || !g) return fail(MPK_EINVAL, "theta, g and tau are required");
    if ((in_dtype != MPK_F64 && in_dtype != MPK_F32) || (out_dtype != MPK_F64 && out_dtype != MPK_F32))
        return fail(MPK_EINVAL, "bad dtype");
    RneaArgs a;
    a.P = P;
    a.th = theta;
    a.dth = dtheta;
    a.ddth = ddtheta;
    a.in_dtype = in_dtype;
    a.vec_in = aligned16(theta) && (!dtheta || aligned16(dtheta)) && (!ddtheta || aligned16(ddtheta));
    a.tip = make_tip(rb, g, Ftip, Ftip_rows);
    a.lim = make_limits(out_dtype == MPK_F32 ? tau_limits : nullptr, rb->n);
    a.out = tau;
    a.out_dtype = out_dtype;
    a.vec_out = aligned16(tau);
    a.compute_f32 = compute_f32;
    unsigned grid;
    if (int rc = grid_for(P, grid)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MPK_DISPATCH_FLAVOUR(rb, launch_rnea<F_>(rb, a, grid, s));
    return check_launch("inverse_dynamics");
}

extern "C" int mpk_inverse_dynamics(const mpk_robot *rb, int64_t P, const void *theta,
                                    const void *dtheta, const void *ddtheta, int in_dtype,
                                    const double *g, const double *Ftip, const double *Ftip_rows,
                                    const float *tau_limits, void *tau, int out_dtype,
                                    void *stream) {
    return inverse_dynamics_impl(rb, P, theta, dtheta, ddtheta, in_dtype, g, Ftip, Ftip_rows, tau_limits, tau,
                                 out_dtype, 0, stream);
}

extern "C" int mpk_inverse_dynamics_f32(const mpk_robot *rb, int64_t P, const void *theta,
                                        const void *dtheta, const void *ddtheta, int in_dtype,
                                        const double *g, const double *Ftip, const double *Ftip_rows,
                                        const float *tau_limits, void *tau, int out_dtype,
                                        void *stream) {
    return inverse_dynamics_impl(rb, P, theta, dtheta, ddtheta, in_dtype, g, Ftip, Ftip_rows, tau_limits, tau,
                                 out_dtype, 1, stream);
}

static int trajectory_inverse_dynamics_impl(const mpk_robot *rb, int64_t B, int64_t N,
                                            const double *start, const double *end,
                                            int inputs_f32, double Tf, int method,
                                            const float *joint_limits, const double *g,
                                            const double *Ftip, const float *tau_limits,
                                            float *tau, float *pos, float *vel, float *acc,
                                            double *ts_scratch, int compute_f32, void *stream) {
    MPK_REQUIRE_DYN(rb);
    if (B < 0 || N < 0) return fail(MPK_EINVAL, "negative size");
    if (B == 0 || N == 0) return MPK_OK;
    if (!start || !end || !tau || !g) return fail(MPK_EINVAL, "start, end, g and tau are required");
    // (tau may be a row-offset view of a larger, e.g. peer-mapped, buffer: any float alignment)
    if (reinterpret_cast<uintptr_t>(tau) & 3u) return fail(MPK_EINVAL, "tau must be 4-byte aligned");
    for (float *o : {pos, vel, acc})
        if (o && !aligned16(o)) return fail(MPK_EINVAL, "trajectory outputs must be 16-byte aligned");
    TrajRneaArgs a;
    a.B = B;
    a.N = N;
    a.P = B * N;
    a.div = make_fastdiv(N, a.P);
    a.start = start;
    a.end = end;
    a.inputs_f32 = inputs_f32;
    a.Tf = Tf;
    a.method = method;
    a.jlim = make_limits(joint_limits, rb->n);
    a.tlim = make_limits(tau_limits, rb->n);
    a.tip = make_tip(rb, g, Ftip, nullptr);
    a.tau = tau;
    a.pos = a.vel = a.acc = nullptr;
    a.compute_f32 = compute_f32;
    a.table_rows = traj_table_rows(B, N);
    unsigned grid;
    if (int rc = grid_for(a.P, grid)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    a.ts_table = prepare_time_scaling(ts_scratch, B, N, Tf, method, s);
    if (pos || vel || acc) {
        // optional materialisation of the trajectory rows
        if (!compute_f32 && !a.tip.has_ftip) {
            // the fused kernel stores them itself, hidden under its arithmetic
            a.pos = pos;
            a.vel = vel;
            a.acc = acc;
        } else if (int rc = launch_joint_trajectory(rb->n, B, N, start, end, inputs_f32, Tf, method,
                                                    joint_limits, pos, vel, acc, a.ts_table, s)) {
            return rc;  // (other variants: the store-bound trajectory kernel writes them first)
        }
    }
    MPK_DISPATCH_FLAVOUR(rb, launch_traj_rnea<F_>(rb, a, grid, s));
    return check_launch("trajectory_inverse_dynamics");
}

extern "C" int mpk_trajectory_inverse_dynamics(const mpk_robot *rb, int64_t B, int64_t N,
                                               const double *start, const double *end,
                                               int inputs_f32, double Tf, int method,
                                               const float *joint_limits, const double *g,
                                               const double *Ftip, const float *tau_limits,
                                               float *tau, float *pos, float *vel, float *acc,
                                               double *ts_scratch, void *stream) {
    return trajectory_inverse_dynamics_impl(rb, B, N, start, end, inputs_f32, Tf, method, joint_limits, g, Ftip,
                                            tau_limits, tau, pos, vel, acc, ts_scratch, 0, stream);
}

extern "C" int mpk_trajectory_inverse_dynamics_f32(const mpk_robot *rb, int64_t B, int64_t N,
                                                   const double *start, const double *end,
                                                   int inputs_f32, double Tf, int method,
                                                   const float *joint_limits, const double *g,
                                                   const double *Ftip, const float *tau_limits,
                                                   float *tau, float *pos, float *vel, float *acc,
                                                   double *ts_scratch, void *stream) {
    return trajectory_inverse_dynamics_impl(rb, B, N, start, end, inputs_f32, Tf, method, joint_limits, g, Ftip,
                                            tau_limits, tau, pos, vel, acc, ts_scratch, 1, stream);
}

extern "C" int mpk_mass_matrix(const mpk_robot *rb, int64_t P, const void *theta, int theta_dtype,
                               double *Mout, void *stream) {
    MPK_REQUIRE_DYN(rb);
    if (P < 0) return fail(MPK_EINVAL, "negative size");
    if (P == 0) return MPK_OK;
    if (!theta || !Mout) return fail(MPK_EINVAL, "theta and Mout are required");
    if (theta_dtype != MPK_F64 && theta_dtype != MPK_F32) return fail(MPK_EINVAL, "bad dtype");
    MassArgs a;
    a.P = P;
    a.th = theta;
    a.th_dtype = theta_dtype;
    a.vec_in = aligned16(theta);
    a.vec_out = aligned16(Mout);
    a.out = Mout;
    unsigned grid;
    if (int rc = grid_for(P, grid)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MPK_DISPATCH_FLAVOUR(rb, launch_mass<F_>(rb, a, grid, s));
    return check_launch("mass_matrix");
}

extern "C" int mpk_forward_dynamics(const mpk_robot *rb, int64_t P, const double *theta,
                                    const double *dtheta, const double *tau, const double *g,
                                    const double *Ftip, const double *Ftip_rows, double *ddtheta,
                                    void *stream) {
    MPK_REQUIRE_DYN(rb);
    if (P < 0) return fail(MPK_EINVAL, "negative size");
    if (P == 0) return MPK_OK;
    if (!theta || !dtheta || !tau || !g || !ddtheta)
        return fail(MPK_EINVAL, "theta, dtheta, tau, g and ddtheta are required");
    FdArgs a;
    a.P = P;
    a.th = theta;
    a.dth = dtheta;
    a.tau = tau;
    a.vec = aligned16(theta) && aligned16(dtheta) && aligned16(tau) && aligned16(ddtheta);
    a.tip = make_tip(rb, g, Ftip, Ftip_rows);
    a.out = ddtheta;
    unsigned grid;
    if (int rc = grid_for(P, grid)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MPK_DISPATCH_FLAVOUR(rb, launch_fd_point<F_>(rb, a, grid, s));
    return check_launch("forward_dynamics");
}

extern "C" int mpk_forward_dynamics_trajectory(const mpk_robot *rb, int64_t B, int64_t N,
                                               const double *theta0, const double *dtheta0,
                                               const void *taumat, int tau_dtype, const double *g,
                                               const double *Ftipmat, double dt, int intRes,
                                               const float *limits, float *pos, float *vel,
                                               float *acc, void *stream) {
    MPK_REQUIRE_DYN(rb);
    if (B < 0 || N < 0 || intRes < 1) return fail(MPK_EINVAL, "bad sizes");
    if (B == 0 || N == 0) return MPK_OK;
    if (!theta0 || !dtheta0 || !taumat || !g || !pos || !vel || !acc)
        return fail(MPK_EINVAL, "theta0, dtheta0, taumat, g, pos, vel, acc are required");
    if (tau_dtype != MPK_F64 && tau_dtype != MPK_F32) return fail(MPK_EINVAL, "bad dtype");
    RolloutArgs a;
    a.B = B;
    a.N = N;
    a.th0 = theta0;
    a.dth0 = dtheta0;
    a.taumat = taumat;
    a.tau_dtype = tau_dtype;
    a.vec_tau = aligned16(taumat);
    base_gravity(rb->pack, g, a.g0);
    a.ftipmat = Ftipmat;
    a.dts = dt / (double)intRes;
    a.intRes = intRes;
    a.lim = make_limits(limits, rb->n);
    a.pos = pos;
    a.vel = vel;
    a.acc = acc;
    if ((B + 31) / 32 > 0x7fffffffLL) return fail(MPK_EINVAL, "B exceeds the grid limit");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MPK_DISPATCH_FLAVOUR(rb, launch_rollout<F_>(rb, a, s));
    return check_launch("forward_dynamics_trajectory");
}
