// mpk_torch.cpp -- PyTorch custom-op extension: torch.ops.mpk.* are thin callers of the
// C ABI in include/mpk.h.  PyTorch supplies device memory, the current stream and the
// dispatcher; all arithmetic happens in libmpk.so's hand-written sm_100a kernels.  There
// is no CPU implementation: every op requires CUDA tensors and raises otherwise.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <tuple>
#include <vector>

#include "mpk.h"

namespace {

using at::Tensor;
using OptT = std::optional<Tensor>;

void check(int rc, const char *what) {
    TORCH_CHECK(rc == MPK_OK, "mpk::", what, " failed (", rc, "): ", mpk_last_error());
}

mpk_robot *robot(int64_t h) {
    TORCH_CHECK(h != 0, "mpk: null robot handle");
    return reinterpret_cast<mpk_robot *>(static_cast<intptr_t>(h));
}

void *stream_of(const Tensor &t) { return c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

Tensor dev_rows(const Tensor &t, int64_t n, const char *name, bool allow_f32) {
    TORCH_CHECK(t.is_cuda(), "mpk: ", name, " must be a CUDA tensor (there is no CPU path)");
    TORCH_CHECK(t.scalar_type() == at::kDouble || (allow_f32 && t.scalar_type() == at::kFloat),
                "mpk: ", name, " must be float64", allow_f32 ? " or float32" : "");
    TORCH_CHECK(t.dim() >= 1 && t.size(-1) == n, "mpk: ", name, " must have ", n, " columns");
    return t.contiguous();
}
int dtype_of(const Tensor &t) { return t.scalar_type() == at::kDouble ? MPK_F64 : MPK_F32; }

// small host constants -------------------------------------------------------------------
std::vector<float> host_limits(const OptT &lim, int64_t n, const char *name) {
    std::vector<float> v;
    if (!lim.has_value()) return v;
    Tensor l = lim->to(at::kCPU, at::kFloat).contiguous();
    TORCH_CHECK(l.numel() == 2 * n, "mpk: ", name, " must be (", n, ", 2)");
    v.assign(l.data_ptr<float>(), l.data_ptr<float>() + 2 * n);
    return v;
}
const float *ptr_or_null(const std::vector<float> &v) { return v.empty() ? nullptr : v.data(); }

std::vector<double> host_vec(c10::ArrayRef<double> a, size_t n, const char *name) {
    TORCH_CHECK(a.size() == n, "mpk: ", name, " must have ", n, " entries");
    return std::vector<double>(a.begin(), a.end());
}

// ops ---------------------------------------------------------------------------------------
int64_t robot_create(const Tensor &S, const Tensor &M, const OptT &G, const OptT &Mcom, int64_t flags) {
    Tensor s = S.to(at::kCPU, at::kDouble).contiguous();
    Tensor m = M.to(at::kCPU, at::kDouble).contiguous();
    TORCH_CHECK(s.dim() == 2 && s.size(0) == 6, "mpk: S_list must be (6, n)");
    TORCH_CHECK(m.dim() == 2 && m.size(0) == 4 && m.size(1) == 4, "mpk: M must be (4, 4)");
    const int64_t n = s.size(1);
    Tensor g, mc;
    const double *gp = nullptr, *mp = nullptr;
    if (G.has_value()) {
        TORCH_CHECK(Mcom.has_value(), "mpk: Glist needs Mlist_per_link");
        g = G->to(at::kCPU, at::kDouble).contiguous();
        mc = Mcom->to(at::kCPU, at::kDouble).contiguous();
        TORCH_CHECK(g.numel() == n * 36, "mpk: Glist must be (n, 6, 6)");
        TORCH_CHECK(mc.numel() == n * 16, "mpk: Mlist_per_link must be (n, 4, 4)");
        gp = g.data_ptr<double>();
        mp = mc.data_ptr<double>();
    }
    mpk_robot *rb = nullptr;
    check(mpk_robot_create((int)n, s.data_ptr<double>(), m.data_ptr<double>(), gp, mp, (int)flags, &rb),
          "robot_create");
    return static_cast<int64_t>(reinterpret_cast<intptr_t>(rb));
}

void robot_destroy(int64_t h) { mpk_robot_destroy(robot(h)); }
int64_t robot_dof(int64_t h) { return mpk_robot_dof(robot(h)); }
bool robot_is_rigid(int64_t h) { return mpk_robot_is_rigid(robot(h)) == 1; }

std::tuple<Tensor, Tensor, Tensor> joint_trajectory(const Tensor &start, const Tensor &end,
                                                    bool inputs_f32, double Tf, int64_t N,
                                                    int64_t method, const OptT &limits) {
    TORCH_CHECK(start.dim() == 2, "mpk: start must be (B, n)");
    const int64_t B = start.size(0), n = start.size(1);
    Tensor s = dev_rows(start, n, "start", false), e = dev_rows(end, n, "end", false);
    TORCH_CHECK(e.sizes() == s.sizes(), "mpk: start/end shape mismatch");
    TORCH_CHECK(N >= 0, "mpk: N must be >= 0");
    c10::cuda::CUDAGuard guard(s.device());
    auto opt = s.options().dtype(at::kFloat);
    Tensor pos = at::empty({B, N, n}, opt), vel = at::empty({B, N, n}, opt), acc = at::empty({B, N, n}, opt);
    auto lim = host_limits(limits, n, "joint_limits");
    Tensor scratch = at::empty({3, N}, s.options());  // time-scaling table workspace
    check(mpk_joint_trajectory((int)n, B, N, s.data_ptr<double>(), e.data_ptr<double>(), inputs_f32, Tf,
                               (int)method, ptr_or_null(lim), pos.data_ptr<float>(),
                               vel.data_ptr<float>(), acc.data_ptr<float>(), scratch.data_ptr<double>(),
                               stream_of(s)),
          "joint_trajectory");
    return {pos, vel, acc};
}

std::tuple<Tensor, Tensor> fk_jacobian(int64_t h, const Tensor &theta, bool want_T, bool want_J,
                                       bool compute_f32, bool body) {
    mpk_robot *rb = robot(h);
    const int64_t n = mpk_robot_dof(rb);
    Tensor th = dev_rows(theta, n, "theta", true);
    const int64_t P = th.numel() / n;
    c10::cuda::CUDAGuard guard(th.device());
    auto opt = th.options().dtype(compute_f32 ? at::kFloat : at::kDouble);
    Tensor T = want_T ? at::empty({P, 4, 4}, opt) : at::empty({0}, opt);
    Tensor J = want_J ? at::empty({P, 6, n}, opt) : at::empty({0}, opt);
    check(mpk_fk_jacobian(rb, P, th.data_ptr(), dtype_of(th), body ? MPK_FRAME_BODY : MPK_FRAME_SPACE,
                          compute_f32 ? MPK_F32 : MPK_F64, want_T ? T.data_ptr() : nullptr,
                          want_J ? J.data_ptr() : nullptr, stream_of(th)),
          "fk_jacobian");
    return {T, J};
}

Tensor inverse_dynamics(int64_t h, const Tensor &theta, const OptT &dtheta, const OptT &ddtheta,
                        c10::ArrayRef<double> g, std::optional<c10::ArrayRef<double>> ftip,
                        const OptT &ftip_rows, const OptT &tau_limits, bool out_f32, bool compute_f32) {
    mpk_robot *rb = robot(h);
    const int64_t n = mpk_robot_dof(rb);
    Tensor th = dev_rows(theta, n, "theta", true);
    const int64_t P = th.numel() / n;
    Tensor dth, ddth, fr;
    const void *dp = nullptr, *ddp = nullptr;
    const double *frp = nullptr;
    if (dtheta.has_value()) {
        dth = dev_rows(*dtheta, n, "dtheta", true).to(th.scalar_type());
        TORCH_CHECK(dth.numel() == th.numel(), "mpk: dtheta shape mismatch");
        dp = dth.data_ptr();
    }
    if (ddtheta.has_value()) {
        ddth = dev_rows(*ddtheta, n, "ddtheta", true).to(th.scalar_type());
        TORCH_CHECK(ddth.numel() == th.numel(), "mpk: ddtheta shape mismatch");
        ddp = ddth.data_ptr();
    }
    if (ftip_rows.has_value()) {
        fr = dev_rows(*ftip_rows, 6, "Ftip rows", false);
        TORCH_CHECK(fr.numel() == P * 6, "mpk: Ftip rows must be (P, 6)");
        frp = fr.data_ptr<double>();
    }
    auto gv = host_vec(g, 3, "g");
    std::vector<double> fv;
    if (ftip.has_value()) fv = host_vec(*ftip, 6, "Ftip");
    auto lim = host_limits(tau_limits, n, "torque_limits");
    c10::cuda::CUDAGuard guard(th.device());
    Tensor tau = at::empty({P, n}, th.options().dtype(out_f32 ? at::kFloat : at::kDouble));
    check((compute_f32 ? mpk_inverse_dynamics_f32 : mpk_inverse_dynamics)(
              rb, P, th.data_ptr(), dp, ddp, dtype_of(th), gv.data(), fv.empty() ? nullptr : fv.data(), frp,
              ptr_or_null(lim), tau.data_ptr(), out_f32 ? MPK_F32 : MPK_F64, stream_of(th)),
          "inverse_dynamics");
    return tau;
}

std::tuple<Tensor, Tensor, Tensor, Tensor> trajectory_inverse_dynamics(
    int64_t h, const Tensor &start, const Tensor &end, bool inputs_f32, double Tf, int64_t N,
    int64_t method, const OptT &joint_limits, c10::ArrayRef<double> g,
    std::optional<c10::ArrayRef<double>> ftip, const OptT &tau_limits, bool want_traj, bool compute_f32,
    const OptT &out) {
    mpk_robot *rb = robot(h);
    const int64_t n = mpk_robot_dof(rb);
    TORCH_CHECK(start.dim() == 2, "mpk: start must be (B, n)");
    Tensor s = dev_rows(start, n, "start", false), e = dev_rows(end, n, "end", false);
    TORCH_CHECK(e.sizes() == s.sizes(), "mpk: start/end shape mismatch");
    TORCH_CHECK(N >= 0, "mpk: N must be >= 0");
    const int64_t B = s.size(0);
    auto gv = host_vec(g, 3, "g");
    std::vector<double> fv;
    if (ftip.has_value()) fv = host_vec(*ftip, 6, "Ftip");
    auto jl = host_limits(joint_limits, n, "joint_limits");
    auto tl = host_limits(tau_limits, n, "torque_limits");
    c10::cuda::CUDAGuard guard(s.device());
    auto opt = s.options().dtype(at::kFloat);
    // `out`: caller-owned float32 (B, N, n) destination.  It may live in ANOTHER GPU's memory (a
    // peer-mapped buffer of the rank that collects the result): the kernel's coalesced row stores then
    // travel over NVLink as they are produced -- the gather is fused into the kernel, tile by tile.
    Tensor tau;
    if (out.has_value()) {
        TORCH_CHECK(out->is_cuda() && out->scalar_type() == at::kFloat && out->is_contiguous() &&
                        out->numel() == B * N * n,
                    "mpk: out must be a contiguous CUDA float32 tensor of B * N * n elements");
        tau = *out;
    } else {
        tau = at::empty({B, N, n}, opt);
    }
    Tensor pos, vel, acc;
    float *pp = nullptr, *vp = nullptr, *ap = nullptr;
    if (want_traj) {
        pos = at::empty({B, N, n}, opt);
        vel = at::empty({B, N, n}, opt);
        acc = at::empty({B, N, n}, opt);
        pp = pos.data_ptr<float>();
        vp = vel.data_ptr<float>();
        ap = acc.data_ptr<float>();
    } else {
        pos = vel = acc = at::empty({0}, opt);
    }
    Tensor scratch = at::empty({3, N}, s.options());  // time-scaling table workspace
    check((compute_f32 ? mpk_trajectory_inverse_dynamics_f32 : mpk_trajectory_inverse_dynamics)(
              rb, B, N, s.data_ptr<double>(), e.data_ptr<double>(), inputs_f32, Tf, (int)method, ptr_or_null(jl),
              gv.data(), fv.empty() ? nullptr : fv.data(), ptr_or_null(tl), tau.data_ptr<float>(), pp, vp, ap,
              scratch.data_ptr<double>(), stream_of(s)),
          "trajectory_inverse_dynamics");
    if (out.has_value()) tau = at::empty({0}, opt);  // (the schema declares no aliasing: the caller keeps `out`)
    return {tau, pos, vel, acc};
}

Tensor mass_matrix(int64_t h, const Tensor &theta) {
    mpk_robot *rb = robot(h);
    const int64_t n = mpk_robot_dof(rb);
    Tensor th = dev_rows(theta, n, "theta", true);
    const int64_t P = th.numel() / n;
    c10::cuda::CUDAGuard guard(th.device());
    Tensor M = at::empty({P, n, n}, th.options().dtype(at::kDouble));
    check(mpk_mass_matrix(rb, P, th.data_ptr(), dtype_of(th), M.data_ptr<double>(), stream_of(th)),
          "mass_matrix");
    return M;
}

Tensor forward_dynamics(int64_t h, const Tensor &theta, const Tensor &dtheta, const Tensor &tau,
                        c10::ArrayRef<double> g, std::optional<c10::ArrayRef<double>> ftip,
                        const OptT &ftip_rows) {
    mpk_robot *rb = robot(h);
    const int64_t n = mpk_robot_dof(rb);
    Tensor th = dev_rows(theta, n, "theta", false), dth = dev_rows(dtheta, n, "dtheta", false),
           ta = dev_rows(tau, n, "tau", false);
    const int64_t P = th.numel() / n;
    TORCH_CHECK(dth.numel() == th.numel() && ta.numel() == th.numel(), "mpk: shape mismatch");
    Tensor fr;
    const double *frp = nullptr;
    if (ftip_rows.has_value()) {
        fr = dev_rows(*ftip_rows, 6, "Ftip rows", false);
        TORCH_CHECK(fr.numel() == P * 6, "mpk: Ftip rows must be (P, 6)");
        frp = fr.data_ptr<double>();
    }
    auto gv = host_vec(g, 3, "g");
    std::vector<double> fv;
    if (ftip.has_value()) fv = host_vec(*ftip, 6, "Ftip");
    c10::cuda::CUDAGuard guard(th.device());
    Tensor dd = at::empty({P, n}, th.options());
    check(mpk_forward_dynamics(rb, P, th.data_ptr<double>(), dth.data_ptr<double>(), ta.data_ptr<double>(),
                               gv.data(), fv.empty() ? nullptr : fv.data(), frp, dd.data_ptr<double>(),
                               stream_of(th)),
          "forward_dynamics");
    return dd;
}

std::tuple<Tensor, Tensor, Tensor> forward_dynamics_trajectory(
    int64_t h, const Tensor &theta0, const Tensor &dtheta0, const Tensor &taumat, c10::ArrayRef<double> g,
    const OptT &ftipmat, double dt, int64_t intRes, const OptT &limits) {
    mpk_robot *rb = robot(h);
    const int64_t n = mpk_robot_dof(rb);
    TORCH_CHECK(theta0.dim() == 2 && taumat.dim() == 3, "mpk: theta0 (B, n), taumat (B, N, n)");
    Tensor th = dev_rows(theta0, n, "theta0", false), dth = dev_rows(dtheta0, n, "dtheta0", false);
    Tensor tm = dev_rows(taumat, n, "taumat", true);
    const int64_t B = th.size(0), N = tm.size(1);
    TORCH_CHECK(dth.sizes() == th.sizes() && tm.size(0) == B, "mpk: batch mismatch");
    TORCH_CHECK(intRes >= 1, "mpk: intRes must be >= 1");
    Tensor fm;
    const double *fmp = nullptr;
    if (ftipmat.has_value()) {
        fm = dev_rows(*ftipmat, 6, "Ftipmat", false);
        TORCH_CHECK(fm.numel() == B * N * 6, "mpk: Ftipmat must be (B, N, 6)");
        fmp = fm.data_ptr<double>();
    }
    auto gv = host_vec(g, 3, "g");
    auto lim = host_limits(limits, n, "joint_limits");
    c10::cuda::CUDAGuard guard(th.device());
    auto opt = th.options().dtype(at::kFloat);
    Tensor pos = at::empty({B, N, n}, opt), vel = at::empty({B, N, n}, opt), acc = at::empty({B, N, n}, opt);
    check(mpk_forward_dynamics_trajectory(rb, B, N, th.data_ptr<double>(), dth.data_ptr<double>(),
                                          tm.data_ptr(), dtype_of(tm), gv.data(), fmp, dt, (int)intRes,
                                          ptr_or_null(lim), pos.data_ptr<float>(), vel.data_ptr<float>(),
                                          acc.data_ptr<float>(), stream_of(th)),
          "forward_dynamics_trajectory");
    return {pos, vel, acc};
}

std::tuple<Tensor, Tensor, Tensor, Tensor> inverse_kinematics_dls(
    int64_t h, const Tensor &T_desired, const Tensor &theta0, double eomg, double ev, int64_t max_iterations,
    double damping, double step_cap, double weight_orientation, double weight_position, const OptT &limits,
    int64_t seed, bool two_phase, int64_t flags, const OptT &restart_noise) {
    mpk_robot *rb = robot(h);
    const int64_t n = mpk_robot_dof(rb);
    Tensor th0 = dev_rows(theta0, n, "thetalist0", false);
    const int64_t P = th0.numel() / n;
    TORCH_CHECK(T_desired.is_cuda() && T_desired.numel() == P * 16, "mpk: T_desired must be a CUDA (P, 4, 4) tensor");
    Tensor Td = T_desired.to(at::kDouble).contiguous();
    std::vector<double> lim;
    if (limits.has_value()) {
        Tensor l = limits->to(at::kCPU, at::kDouble).contiguous();
        TORCH_CHECK(l.numel() == n * 2, "mpk: joint_limits must be (n, 2)");
        lim.assign(l.data_ptr<double>(), l.data_ptr<double>() + n * 2);
    }
    c10::cuda::CUDAGuard guard(th0.device());
    Tensor theta = at::empty({P, n}, th0.options());
    Tensor iters = at::empty({P}, th0.options().dtype(at::kInt));
    Tensor ok = at::empty({P}, th0.options().dtype(at::kByte));
    Tensor restarts = at::empty({P}, th0.options().dtype(at::kInt));
    Tensor noise;
    int noise_rows = 0;
    if (restart_noise.has_value() && P > 0) {
        TORCH_CHECK(restart_noise->is_cuda() && restart_noise->numel() % (P * n) == 0,
                    "mpk: restart_noise must be a CUDA (P, rows, n) tensor");
        noise = restart_noise->to(at::kDouble).contiguous();
        noise_rows = (int)(noise.numel() / (P * n));
    }
    const size_t ws_bytes = two_phase ? mpk_inverse_kinematics_workspace_bytes((int)n, P) : 0;
    Tensor ws = at::empty({(int64_t)(ws_bytes / 8) + 1}, th0.options());
    check(mpk_inverse_kinematics_dls_modes(rb, P, Td.data_ptr<double>(), th0.data_ptr<double>(), eomg, ev,
                                           (int)max_iterations, damping, step_cap, weight_orientation,
                                           weight_position, lim.empty() ? nullptr : lim.data(), (int)flags,
                                           (uint64_t)seed, noise_rows ? noise.data_ptr<double>() : nullptr, noise_rows,
                                           theta.data_ptr<double>(), iters.data_ptr<int32_t>(),
                                           ok.data_ptr<uint8_t>(), restarts.data_ptr<int32_t>(),
                                           two_phase ? ws.data_ptr() : nullptr, ws_bytes, stream_of(th0)),
          "inverse_kinematics_dls");
    return {theta, ok, iters, restarts};
}

std::tuple<Tensor, Tensor, Tensor, Tensor> cartesian_trajectory(const Tensor &Xstart, const Tensor &Xend, double Tf,
                                                                int64_t N, int64_t method) {
    TORCH_CHECK(Xstart.is_cuda() && Xend.is_cuda(), "mpk: Xstart / Xend must be CUDA tensors");
    Tensor xs = Xstart.to(at::kDouble).contiguous(), xe = Xend.to(at::kDouble).contiguous();
    TORCH_CHECK(xs.numel() % 16 == 0 && xe.numel() == xs.numel(), "mpk: Xstart / Xend must be (B, 4, 4)");
    TORCH_CHECK(N >= 0, "mpk: N must be >= 0");
    const int64_t B = xs.numel() / 16;
    c10::cuda::CUDAGuard guard(xs.device());
    auto opt = xs.options().dtype(at::kFloat);
    Tensor pos = at::empty({B, N, 3}, opt), vel = at::empty({B, N, 3}, opt), acc = at::empty({B, N, 3}, opt);
    Tensor ori = at::empty({B, N, 3, 3}, opt);
    check(mpk_cartesian_trajectory(B, N, xs.data_ptr<double>(), xe.data_ptr<double>(), Tf, (int)method,
                                   pos.data_ptr<float>(), vel.data_ptr<float>(), acc.data_ptr<float>(),
                                   ori.data_ptr<float>(), stream_of(xs)),
          "cartesian_trajectory");
    return {pos, vel, acc, ori};
}

void fma_peak(const Tensor &sink, int64_t dtype, int64_t blocks, int64_t threads, int64_t iters) {
    TORCH_CHECK(sink.is_cuda() && sink.scalar_type() == at::kDouble && sink.numel() >= 1,
                "mpk: sink must be a CUDA float64 tensor");
    c10::cuda::CUDAGuard guard(sink.device());
    check(mpk_fma_peak((int)dtype, (int)blocks, (int)threads, iters, sink.data_ptr<double>(), stream_of(sink)),
          "fma_peak");
}

// ---- legacy dynamics path (csrc/legacy.cu) -----------------------------------------------------------
Tensor legacy_dynamics(const Tensor &S_list, const Tensor &M, const Tensor &Glist, int64_t mode, const Tensor &theta,
                       const OptT &dtheta, const OptT &third, c10::ArrayRef<double> g,
                       std::optional<c10::ArrayRef<double>> ftip, const OptT &ftip_rows) {
    Tensor S = S_list.to(at::kCPU, at::kDouble).contiguous(), Mh = M.to(at::kCPU, at::kDouble).contiguous();
    Tensor G = Glist.to(at::kCPU, at::kDouble).contiguous();
    TORCH_CHECK(S.dim() == 2 && S.size(0) == 6 && Mh.numel() == 16, "mpk: S_list must be (6, n) and M (4, 4)");
    const int64_t n = S.size(1);
    TORCH_CHECK(G.numel() == n * 36, "mpk: Glist must be (n, 6, 6)");
    Tensor th = dev_rows(theta, n, "theta", false);
    const int64_t P = th.numel() / n;
    Tensor d1, d3, fr;
    const double *p1 = nullptr, *p3 = nullptr, *pf = nullptr;
    if (dtheta.has_value()) { d1 = dev_rows(*dtheta, n, "dtheta", false); p1 = d1.data_ptr<double>(); }
    if (third.has_value()) { d3 = dev_rows(*third, n, "ddtheta / tau", false); p3 = d3.data_ptr<double>(); }
    if (ftip_rows.has_value()) { fr = dev_rows(*ftip_rows, 6, "Ftip rows", false); pf = fr.data_ptr<double>(); }
    auto gv = host_vec(g, 3, "g");
    std::vector<double> fv;
    if (ftip.has_value()) fv = host_vec(*ftip, 6, "Ftip");
    c10::cuda::CUDAGuard guard(th.device());
    Tensor out = mode == 0 ? at::empty({P, n, n}, th.options()) : at::empty({P, n}, th.options());
    check(mpk_legacy_dynamics((int)n, S.data_ptr<double>(), Mh.data_ptr<double>(), G.data_ptr<double>(), (int)mode, P,
                              th.data_ptr<double>(), p1, p3, gv.data(), fv.empty() ? nullptr : fv.data(), pf,
                              out.data_ptr<double>(), stream_of(th)),
          "legacy_dynamics");
    return out;
}

// ---- collision hook (csrc/collision.cu) ---------------------------------------------------------
Tensor collision_model_pack(int64_t h, const Tensor &link_joint, const Tensor &link_home, const Tensor &acm,
                            const Tensor &hull_link, const Tensor &hull_count, const Tensor &hull_points) {
    mpk_robot *rb = robot(h);
    Tensor lj = link_joint.to(at::kCPU, at::kInt).contiguous(), lh = link_home.to(at::kCPU, at::kDouble).contiguous();
    Tensor ac = acm.to(at::kCPU, at::kByte).contiguous();
    Tensor hl = hull_link.to(at::kCPU, at::kInt).contiguous(), hc = hull_count.to(at::kCPU, at::kInt).contiguous();
    Tensor hp = hull_points.to(at::kCPU, at::kDouble).contiguous();
    const int64_t L = lj.numel(), H = hl.numel();
    TORCH_CHECK(lh.numel() == L * 16 && ac.numel() == L * L, "mpk: link_home must be (L, 4, 4) and acm (L, L)");
    TORCH_CHECK(hc.numel() == H && hp.numel() == hc.sum().item<int64_t>() * 3, "mpk: hull tables do not match");
    const size_t bytes = mpk_collision_model_bytes(mpk_robot_dof(rb), (int)L, (int)H, hp.numel() / 3);
    Tensor out = at::zeros({(int64_t)bytes}, at::TensorOptions().dtype(at::kByte));
    check(mpk_collision_model_pack(rb, (int)L, lj.data_ptr<int32_t>(), lh.data_ptr<double>(), ac.data_ptr<uint8_t>(),
                                   (int)H, H ? hl.data_ptr<int32_t>() : nullptr, H ? hc.data_ptr<int32_t>() : nullptr,
                                   H ? hp.data_ptr<double>() : nullptr, out.data_ptr<uint8_t>(), bytes),
          "collision_model_pack");
    return out;
}

Tensor link_fk_batch(int64_t h, const Tensor &model_host, const Tensor &model_dev, const Tensor &theta, int64_t L) {
    mpk_robot *rb = robot(h);
    const int64_t n = mpk_robot_dof(rb);
    Tensor th = dev_rows(theta, n, "theta", true);
    TORCH_CHECK(model_dev.is_cuda() && !model_host.is_cuda(), "mpk: model_host / model_dev");
    const int64_t P = th.numel() / n;
    c10::cuda::CUDAGuard guard(th.device());
    Tensor T = at::empty({P, L, 4, 4}, th.options().dtype(at::kDouble));
    check(mpk_link_fk_batch(rb, model_host.data_ptr(), model_dev.data_ptr(), P, th.data_ptr(), dtype_of(th),
                            T.data_ptr<double>(), stream_of(th)),
          "link_fk_batch");
    return T;
}

Tensor self_collision(int64_t h, const Tensor &model_host, const Tensor &model_dev, const Tensor &theta) {
    mpk_robot *rb = robot(h);
    const int64_t n = mpk_robot_dof(rb);
    Tensor th = dev_rows(theta, n, "theta", true);
    TORCH_CHECK(model_dev.is_cuda() && !model_host.is_cuda(), "mpk: model_host / model_dev");
    const int64_t P = th.numel() / n;
    c10::cuda::CUDAGuard guard(th.device());
    Tensor flags = at::empty({P}, th.options().dtype(at::kByte));
    check(mpk_self_collision_aabb(rb, model_host.data_ptr(), model_dev.data_ptr(), P, th.data_ptr(), dtype_of(th),
                                  flags.data_ptr<uint8_t>(), stream_of(th)),
          "self_collision_aabb");
    return flags;
}

// rows (P, n) float32 are modified in place; -> (iterations int32 (P), still-colliding flags uint8 (P))
std::tuple<Tensor, Tensor> collision_avoidance(int64_t h, const Tensor &model_host, const Tensor &model_dev,
                                               Tensor rows, const Tensor &goal, int64_t rows_per_goal,
                                               double attractive_gain, double step, int64_t max_iterations) {
    mpk_robot *rb = robot(h);
    const int64_t n = mpk_robot_dof(rb);
    TORCH_CHECK(rows.is_cuda() && rows.scalar_type() == at::kFloat && rows.is_contiguous() && rows.size(-1) == n,
                "mpk: rows must be a contiguous CUDA float32 (P, n) tensor");
    Tensor g = goal.to(rows.device(), at::kFloat).contiguous();
    const int64_t P = rows.numel() / n;
    TORCH_CHECK(rows_per_goal >= 1 && g.numel() % n == 0 && (P + rows_per_goal - 1) / rows_per_goal <= g.numel() / n,
                "mpk: goal must hold one row per rows_per_goal rows");
    c10::cuda::CUDAGuard guard(rows.device());
    Tensor it = at::empty({P}, rows.options().dtype(at::kInt)), fl = at::empty({P}, rows.options().dtype(at::kByte));
    check(mpk_collision_avoidance(rb, model_host.data_ptr(), model_dev.data_ptr(), P, rows.data_ptr<float>(),
                                  g.data_ptr<float>(), rows_per_goal, attractive_gain, step, (int)max_iterations,
                                  it.data_ptr<int32_t>(), fl.data_ptr<uint8_t>(), stream_of(rows)),
          "collision_avoidance");
    return {it, fl};
}

// ---- peer-shared result buffers (csrc/peer.cu) ----------------------------------------------
// -> (float32 tensor of `numel` elements on `like`'s device, owning the cudaMalloc'd buffer; 64-byte handle)
std::tuple<Tensor, Tensor> peer_alloc(int64_t numel, const Tensor &like) {
    TORCH_CHECK(like.is_cuda() && numel > 0, "mpk: peer_alloc needs a CUDA device and a positive size");
    c10::cuda::CUDAGuard guard(like.device());
    void *p = nullptr;
    Tensor handle = at::empty({MPK_PEER_HANDLE_BYTES}, at::TensorOptions().dtype(at::kByte));
    check(mpk_peer_alloc((size_t)numel * 4, &p, handle.data_ptr<uint8_t>()), "peer_alloc");
    const int dev = like.get_device();
    Tensor t = at::from_blob(
        p, {numel},
        [dev](void *q) {
            c10::cuda::CUDAGuard g((c10::DeviceIndex)dev);
            mpk_peer_free(q);
        },
        at::TensorOptions().dtype(at::kFloat).device(like.device()), like.device());
    return {t, handle};
}

// Map another rank's buffer: a float32 tensor labelled with `like`'s device (the pointer is valid in
// this device's address space; the bytes live in the exporting GPU's HBM).
Tensor peer_open(const Tensor &handle, int64_t numel, const Tensor &like) {
    TORCH_CHECK(like.is_cuda() && numel > 0, "mpk: peer_open needs a CUDA device and a positive size");
    Tensor h = handle.to(at::kCPU, at::kByte).contiguous();
    TORCH_CHECK(h.numel() == MPK_PEER_HANDLE_BYTES, "mpk: handle must be ", MPK_PEER_HANDLE_BYTES, " bytes");
    c10::cuda::CUDAGuard guard(like.device());
    void *p = nullptr;
    check(mpk_peer_open(h.data_ptr<uint8_t>(), &p), "peer_open");
    const int dev = like.get_device();
    return at::from_blob(
        p, {numel},
        [dev](void *q) {
            c10::cuda::CUDAGuard g((c10::DeviceIndex)dev);
            mpk_peer_close(q);
        },
        at::TensorOptions().dtype(at::kFloat).device(like.device()), like.device());
}

void store_peak(const Tensor &dst, int64_t mode, int64_t blocks) {
    TORCH_CHECK(dst.is_cuda() && dst.is_contiguous(), "mpk: dst must be a contiguous CUDA tensor");
    c10::cuda::CUDAGuard guard(dst.device());
    check(mpk_store_peak(dst.data_ptr(), (int64_t)dst.nbytes(), (int)mode, (int)blocks, stream_of(dst)), "store_peak");
}

}  // namespace

TORCH_LIBRARY(mpk, m) {
    m.def("robot_create(Tensor S_list, Tensor M, Tensor? Glist, Tensor? Mlist_per_link, int flags) -> int",
          &robot_create);
    m.def("robot_destroy(int robot) -> ()", &robot_destroy);
    m.def("robot_dof(int robot) -> int", &robot_dof);
    m.def("robot_is_rigid(int robot) -> bool", &robot_is_rigid);
    m.def("joint_trajectory(Tensor start, Tensor end, bool inputs_f32, float Tf, int N, int method, "
          "Tensor? joint_limits) -> (Tensor, Tensor, Tensor)",
          &joint_trajectory);
    m.def("fk_jacobian(int robot, Tensor theta, bool want_T, bool want_J, bool compute_f32=False, "
          "bool body=False) -> (Tensor, Tensor)",
          &fk_jacobian);
    m.def("inverse_dynamics(int robot, Tensor theta, Tensor? dtheta, Tensor? ddtheta, float[] g, "
          "float[]? Ftip, Tensor? Ftip_rows, Tensor? torque_limits, bool out_f32, bool compute_f32=False) -> Tensor",
          &inverse_dynamics);
    m.def("trajectory_inverse_dynamics(int robot, Tensor start, Tensor end, bool inputs_f32, float Tf, "
          "int N, int method, Tensor? joint_limits, float[] g, float[]? Ftip, Tensor? torque_limits, "
          "bool want_traj, bool compute_f32=False, Tensor? out=None) -> (Tensor, Tensor, Tensor, Tensor)",
          &trajectory_inverse_dynamics);
    m.def("mass_matrix(int robot, Tensor theta) -> Tensor", &mass_matrix);
    m.def("forward_dynamics(int robot, Tensor theta, Tensor dtheta, Tensor tau, float[] g, float[]? Ftip, "
          "Tensor? Ftip_rows) -> Tensor",
          &forward_dynamics);
    m.def("forward_dynamics_trajectory(int robot, Tensor theta0, Tensor dtheta0, Tensor taumat, float[] g, "
          "Tensor? Ftipmat, float dt, int intRes, Tensor? joint_limits) -> (Tensor, Tensor, Tensor)",
          &forward_dynamics_trajectory);
    m.def("inverse_kinematics_dls(int robot, Tensor T_desired, Tensor theta0, float eomg, float ev, "
          "int max_iterations, float damping, float step_cap, float weight_orientation, float weight_position, "
          "Tensor? joint_limits, int seed, bool two_phase=True, int flags=0, Tensor? restart_noise=None) -> "
          "(Tensor, Tensor, Tensor, Tensor)",
          &inverse_kinematics_dls);
    m.def("cartesian_trajectory(Tensor Xstart, Tensor Xend, float Tf, int N, int method) -> "
          "(Tensor, Tensor, Tensor, Tensor)",
          &cartesian_trajectory);
    m.def("fma_peak(Tensor sink, int dtype, int blocks, int threads, int iters) -> ()", &fma_peak);
    m.def("store_peak(Tensor dst, int mode, int blocks) -> ()", &store_peak);
    m.def("legacy_dynamics(Tensor S_list, Tensor M, Tensor Glist, int mode, Tensor theta, Tensor? dtheta, Tensor? third, "
          "float[] g, float[]? Ftip, Tensor? Ftip_rows) -> Tensor",
          &legacy_dynamics);
    m.def("collision_model_pack(int robot, Tensor link_joint, Tensor link_home, Tensor acm, Tensor hull_link, "
          "Tensor hull_count, Tensor hull_points) -> Tensor",
          &collision_model_pack);
    m.def("link_fk_batch(int robot, Tensor model_host, Tensor model_dev, Tensor theta, int L) -> Tensor", &link_fk_batch);
    m.def("self_collision(int robot, Tensor model_host, Tensor model_dev, Tensor theta) -> Tensor", &self_collision);
    m.def("collision_avoidance(int robot, Tensor model_host, Tensor model_dev, Tensor(a!) rows, Tensor goal, "
          "int rows_per_goal, float attractive_gain, float step, int max_iterations) -> (Tensor, Tensor)",
          &collision_avoidance);
    m.def("peer_alloc(int numel, Tensor like) -> (Tensor, Tensor)", &peer_alloc);
    m.def("peer_open(Tensor handle, int numel, Tensor like) -> Tensor", &peer_open);
}
