/*
 * mpk.h -- C ABI of the B200-native batched trajectory-and-dynamics path.
 *
 * This is the drop-in boundary.  ManipulaPy has no FFI of its
 * own: the seams are its Python methods, so every entry point below names the
 * reference method (file:line, relative to the ManipulaPy v1.4.1 checkout) it
 * replaces.  The Python host (manipulapy_b200/) and the PyTorch custom ops
 * (torch.ops.mpk.*) are thin callers of exactly these symbols; a maintainer of
 * the reference binds them with the ctypes stubs shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in signatures (a stream is
 *     passed as `void*` holding a cudaStream_t; NULL = legacy default stream).
 *   - "dev" pointers are device memory on the CURRENT CUDA device, row-major,
 *     densely packed; "host" pointers are small host arrays read before the
 *     call returns.  No hidden allocation, no synchronisation: every launcher
 *     enqueues on `stream` and returns.
 *   - return value: MPK_OK (0) or a negative MPK_E* code; nothing throws.
 *     mpk_last_error() returns a human-readable reason for the last failure
 *     on the calling thread.
 *   - twists are [omega; v], wrenches [moment; force], S_list is (6, n) with
 *     one column per joint -- the reference's conventions (utils/se3.py:45-52).
 *   - element types: MPK_F64 / MPK_F32 select the STORAGE type of an array;
 *     arithmetic is float64 throughout unless a launcher says otherwise.
 */
#ifndef MPK_H
#define MPK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPK_OK 0
#define MPK_EINVAL (-1)      /* bad argument (NULL pointer, negative size, bad dtype)   */
#define MPK_EUNSUPPORTED (-2) /* robot outside the supported class (see mpk_robot_create) */
#define MPK_ECUDA (-3)       /* CUDA launch/runtime error                                */

#define MPK_F64 0
#define MPK_F32 1

#define MPK_MAX_DOF 8

/* or-ed into `method` of the trajectory launchers: the contract of the reference's registry
 * launchers and CUDA kernels (cuda_kernels/trajectory_kernels.py:40-76, 179, 195-198) instead
 * of the planner's: linear time scaling for method not in {3, 5}, and zero scaling ("sit at
 * start") for N <= 1 or Tf <= 0. */
#define MPK_TRAJ_REGISTRY_CONTRACT 0x100

/* mpk_robot_create flags */
#define MPK_ROBOT_FORCE_GENERAL 1 /* route a rigid-body robot through the general-inertia kernels (tests) */

typedef struct mpk_robot mpk_robot; /* opaque; host-resident constant pack (< 4 KB) */

int mpk_version(void);
const char *mpk_last_error(void);

/* Constant pack of ManipulatorDynamics / SerialManipulator
 * (dynamics/manipulator_dynamics.py:46-75; built by urdf/core.py:670-769).
 *   S_list (6, n)   space screws, unit omega or omega = 0 (utils/se3.py:33-42 assumes this)
 *   M      (4, 4)   end-effector home pose
 *   Glist  (n,6,6)  spatial inertias in the link-CoM frames (NULL: kinematics only)
 *   Mcom   (n,4,4)  Mlist_per_link (NULL with Glist NULL)
 * All host, float64.  1 <= n <= MPK_MAX_DOF.  The pack is re-expressed in
 * joint-aligned (Denavit-Hartenberg / Hayati) link frames on the host and later
 * passed to every kernel as a __grid_constant__ parameter (constant bank), so
 * there is no device allocation. */
int mpk_robot_create(int n, const double *S_list, const double *M, const double *Glist,
                     const double *Mcom, int flags, mpk_robot **out);
void mpk_robot_destroy(mpk_robot *rb);
int mpk_robot_dof(const mpk_robot *rb);
/* 1 if every link inertia is a rigid body expressed at its centre of mass
 * (block-diagonal [I, m*1]); 0 if the general symmetric-6x6 kernels are used. */
int mpk_robot_is_rigid(const mpk_robot *rb);
/* The joint-aligned link frames the pack was re-expressed in (diagnostics, tests, docs):
 * out (n, 8) float64 rows [a, cos alpha, sin alpha, cos beta, sin beta, phi, d, revolute?] of
 * X_i = Tx(a) Rx(alpha) Ry(beta) Rz(phi) Tz(d), the home pose of frame i in frame i-1 (row 0: the
 * base frame's own offsets, a = 0, alpha = beta = 0). */
int mpk_robot_link_geometry(const mpk_robot *rb, double *out);
/* Link geometry classes of a plain revolute chain, 4 bits per link i >= 1 at bit 4 i: 1 consecutive
 * axes perpendicular (sin alpha == 1), 2 parallel (alpha == 0), 4 a_i == 0, 8 d_i == 0.  The arm
 * families of the reference's robot database (UR, iiwa, CRX, LR Mate / M-16, IRB 2400, Gen3) have
 * kernels compiled for their signature; every other robot runs the general code (same results). */
unsigned mpk_robot_geometry_signature(const mpk_robot *rb);
/* 1 if no joint is prismatic (the kernels then drop the prismatic terms at compile time). */
int mpk_robot_all_revolute(const mpk_robot *rb);

/* joint_trajectory / batch_joint_trajectory
 * (planning/trajectory.py:103-169, 276-333, 335-502; kernel :15-75; clip :311-313).
 *   start, end   dev (B, n) float64
 *   inputs_f32   1: round start/end to float32 first and subtract in float32
 *                (joint_trajectory, :147-153); 0: keep float64 (batch path, :474-476)
 *   method       3 cubic, 5 quintic, anything else zero scaling (planner CPU contract :67-68);
 *                | MPK_TRAJ_REGISTRY_CONTRACT for the registry launchers' contract
 *   limits       host (n, 2) float32 joint limits or NULL (no clip)
 *   pos/vel/acc  dev (B, N, n) float32; any may be NULL (not written)
 *   ts_scratch   dev (3, N) float64 workspace or NULL.  The time scaling (s, ds, dds) depends
 *                on the step only, so with a workspace it is evaluated once per step by a
 *                small pre-kernel instead of once per (trajectory, step); results are identical.
 * Time scaling runs in float64 with the reference's operation order and one
 * rounding to float32, so outputs are bit-identical to the reference's. */
int mpk_joint_trajectory(int n, int64_t B, int64_t N, const double *start, const double *end,
                         int inputs_f32, double Tf, int method, const float *limits, float *pos,
                         float *vel, float *acc, double *ts_scratch, void *stream);

/* cartesian_trajectory (planning/trajectory.py:504-594, 676-740) for B start / end pose pairs:
 * orientation Rstart exp(log(Rstart^T Rend) s), position s pend + (1 - s) pstart, linear velocity /
 * acceleration ds (pend - pstart), dds (pend - pstart); s cubic for method 3, quintic otherwise;
 * ds = dds = 0 for a method other than 3 / 5 (the reference's CPU path).  float64 arithmetic,
 * one rounding to float32.
 *   Xstart, Xend dev (B, 4, 4) float64;  pos / vel / acc dev (B, N, 3) float32 or NULL;
 *   orientations dev (B, N, 3, 3) float32 or NULL */
int mpk_cartesian_trajectory(int64_t B, int64_t N, const double *Xstart, const double *Xend, double Tf,
                             int method, float *pos, float *vel, float *acc, float *orientations,
                             void *stream);

/* SerialManipulator.forward_kinematics(theta, "space") (kinematics/fk.py:39-86) and
 * SerialManipulator.jacobian(theta, "space") (kinematics/jacobian.py:39-93), batched.
 *   theta dev (P, n) theta_dtype;  T dev (P, 4, 4) float64 or NULL;  J dev (P, 6, n) float64 or NULL */
int mpk_fk_jacobian_space(const mpk_robot *rb, int64_t P, const void *theta, int theta_dtype,
                          double *T, double *J, void *stream);
/* The same in float32 arithmetic with float32 outputs (north-star tolerance 1e-5 on poses):
 * half the HBM bytes of this store-bound kernel. */
int mpk_fk_jacobian_space_f32(const mpk_robot *rb, int64_t P, const void *theta, int theta_dtype,
                              float *T, float *J, void *stream);

/* General form: frame = MPK_FRAME_SPACE | MPK_FRAME_BODY, out_dtype = MPK_F64 | MPK_F32 selects
 * arithmetic and output type (T, J are arrays of that type).
 * MPK_FRAME_BODY mirrors forward_kinematics(theta, "body") = M prod e^{[B_i] theta_i}
 * (kinematics/fk.py:72-81) and jacobian(theta, "body") (kinematics/jacobian.py:74-90) for a
 * robot created from the screws S'_i = Ad(M) B_i: then T is that pose and J = Ad(T^-1) J_s'
 * is the reference's body Jacobian. */
#define MPK_FRAME_SPACE 0
#define MPK_FRAME_BODY 1
int mpk_fk_jacobian(const mpk_robot *rb, int64_t P, const void *theta, int theta_dtype, int frame,
                    int out_dtype, void *T, void *J, void *stream);

/* SerialManipulator.iterative_inverse_kinematics (kinematics/ik.py:39-311), default mode
 * (adaptive_tuning = backtracking = False), for P independent targets: damped least squares on
 * the space Jacobian with step cap, joint-limit projection, best-iterate tracking and
 * stagnation restart.
 *   T_desired dev (P, 4, 4) float64;  theta0 dev (P, n) float64 (initial guesses, not clipped)
 *   joint_limits host (n, 2) float64, +-inf for an open side, or NULL
 *   seed      key of the counter-based generator used by the stagnation restart (the reference
 *             draws from NumPy's global generator there)
 *   theta dev (P, n) float64;  iterations dev (P) int32 (the reference's k + 1; max_iterations + 1
 *   when exhausted);  success dev (P) uint8 */
int mpk_inverse_kinematics_dls(const mpk_robot *rb, int64_t P, const double *T_desired,
                               const double *theta0, double eomg, double ev, int max_iterations,
                               double damping, double step_cap, double weight_orientation,
                               double weight_position, const double *joint_limits, uint64_t seed,
                               double *theta, int32_t *iterations, uint8_t *success, void *workspace,
                               size_t workspace_bytes, void *stream);
/*   workspace  dev scratch of mpk_inverse_kinematics_workspace_bytes(n, P) bytes, or NULL.  With
 *              it, targets still running after 16, 32, 64, ... iterations are queued and continued by
 *              further, densely packed launches (a warp otherwise lives as long as its slowest
 *              target); results are identical either way. */
size_t mpk_inverse_kinematics_workspace_bytes(int n, int64_t P);

/* The same solver with the reference's optional modes (kinematics/ik.py:215-229, 253-276), which
 * smart_inverse_kinematics / robust_inverse_kinematics (ik.py:327-598) switch on:
 *   MPK_IK_ADAPTIVE_TUNING  Levenberg-Marquardt adaptation of the damping and of the step cap
 *   MPK_IK_BACKTRACKING     line search over the scales 1, 1/2, 1/4, 1/8, 3/4 of the capped step
 * flags = 0 is mpk_inverse_kinematics_dls.
 *   restart_noise  dev (P, noise_rows, n) float64 standard normals, or NULL: restart r of target p adds
 *                  0.1 x row (p, r) to the best iterate (ik.py:206-213).  A caller that draws the rows from
 *                  NumPy's global generator reproduces the reference's restarts draw for draw; beyond
 *                  noise_rows, or with NULL, the counter-based generator keyed by `seed` is used.
 *   restarts       dev (P) int32: restarts taken per target (how many rows were consumed), or NULL */
#define MPK_IK_ADAPTIVE_TUNING 1
#define MPK_IK_BACKTRACKING 2
int mpk_inverse_kinematics_dls_modes(const mpk_robot *rb, int64_t P, const double *T_desired,
                                     const double *theta0, double eomg, double ev, int max_iterations,
                                     double damping, double step_cap, double weight_orientation,
                                     double weight_position, const double *joint_limits, int flags,
                                     uint64_t seed, const double *restart_noise, int noise_rows,
                                     double *theta, int32_t *iterations, uint8_t *success,
                                     int32_t *restarts, void *workspace, size_t workspace_bytes,
                                     void *stream);

/* ManipulatorDynamics.inverse_dynamics (dynamics/id_fd.py:16-48) batched, and
 * inverse_dynamics_trajectory (planning/trajectory_dynamics.py:308-380).
 *   theta/dtheta/ddtheta  dev (P, n) in_dtype; dtheta / ddtheta may be NULL (= 0), which gives
 *                         gravity_forces (dynamics/forces.py:61-133) and
 *                         velocity_quadratic_forces (:26-59, with g = 0)
 *   g             host (3) float64
 *   Ftip          host (6) float64 space-frame wrench for every point, or NULL
 *   Ftip_rows     dev (P, 6) float64 per-point wrench (overrides Ftip), or NULL
 *   tau_limits    host (n, 2) float32 or NULL; applied only when out_dtype = MPK_F32
 *                 (row cast to float32, then clipped: trajectory_dynamics.py:354, 369-373)
 *   tau           dev (P, n) out_dtype */
int mpk_inverse_dynamics(const mpk_robot *rb, int64_t P, const void *theta, const void *dtheta,
                         const void *ddtheta, int in_dtype, const double *g, const double *Ftip,
                         const double *Ftip_rows, const float *tau_limits, void *tau, int out_dtype,
                         void *stream);
/* Same arguments, float32 ARITHMETIC (north-star tolerance 1e-4 relative on torques); the
 * storage types of the arrays are still chosen by in_dtype / out_dtype. */
int mpk_inverse_dynamics_f32(const mpk_robot *rb, int64_t P, const void *theta, const void *dtheta,
                             const void *ddtheta, int in_dtype, const double *g, const double *Ftip,
                             const double *Ftip_rows, const float *tau_limits, void *tau, int out_dtype,
                             void *stream);

/* joint_trajectory + inverse_dynamics_trajectory fused: the trajectory rows are
 * produced in registers, rounded to float32 and clipped exactly as the two-call
 * sequence would, and fed to the inverse dynamics without touching HBM.
 *   start, end dev (B, n) float64;  tau dev (B, N, n) float32
 *   pos/vel/acc dev (B, N, n) float32 or NULL (optional materialisation)
 *   ts_scratch  dev (3, N) float64 workspace or NULL (see mpk_joint_trajectory) */
int mpk_trajectory_inverse_dynamics(const mpk_robot *rb, int64_t B, int64_t N, const double *start,
                                    const double *end, int inputs_f32, double Tf, int method,
                                    const float *joint_limits, const double *g, const double *Ftip,
                                    const float *tau_limits, float *tau, float *pos, float *vel,
                                    float *acc, double *ts_scratch, void *stream);
/* Same arguments; the trajectory rows are produced exactly as above (float64 time scaling, one
 * rounding to float32 -- bit-identical), the inverse dynamics then runs in float32 arithmetic. */
int mpk_trajectory_inverse_dynamics_f32(const mpk_robot *rb, int64_t B, int64_t N, const double *start,
                                        const double *end, int inputs_f32, double Tf, int method,
                                        const float *joint_limits, const double *g, const double *Ftip,
                                        const float *tau_limits, float *tau, float *pos, float *vel,
                                        float *acc, double *ts_scratch, void *stream);

/* ManipulatorDynamics.mass_matrix (dynamics/mass_matrix.py:16-99), batched.
 *   theta dev (P, n) theta_dtype;  Mout dev (P, n, n) float64 (symmetric) */
int mpk_mass_matrix(const mpk_robot *rb, int64_t P, const void *theta, int theta_dtype,
                    double *Mout, void *stream);

/* ManipulatorDynamics.forward_dynamics (dynamics/id_fd.py:50-83), batched.
 *   theta/dtheta/tau dev (P, n) float64; Ftip_rows dev (P, 6) or NULL; ddtheta dev (P, n) float64 */
int mpk_forward_dynamics(const mpk_robot *rb, int64_t P, const double *theta, const double *dtheta,
                         const double *tau, const double *g, const double *Ftip,
                         const double *Ftip_rows, double *ddtheta, void *stream);

/* forward_dynamics_trajectory (planning/trajectory_dynamics.py:580-708) for B
 * independent rollouts (B = 1 is the reference call).
 *   theta0/dtheta0 dev (B, n) float64;  taumat dev (B, N, n) tau_dtype
 *   Ftipmat dev (B, N, 6) float64 or NULL;  limits host (n, 2) float32 or NULL
 *   pos/vel/acc dev (B, N, n) float32: row 0 = initial state with zero acceleration,
 *   row i = state after intRes semi-implicit Euler sub-steps driven by taumat[i]. */
int mpk_forward_dynamics_trajectory(const mpk_robot *rb, int64_t B, int64_t N, const double *theta0,
                                    const double *dtheta0, const void *taumat, int tau_dtype,
                                    const double *g, const double *Ftipmat, double dt, int intRes,
                                    const float *limits, float *pos, float *vel, float *acc,
                                    void *stream);

/* Register-resident FMA micro-benchmark used by bench.py to measure the fp64 /
 * fp32 CUDA-core peak the roofline fractions are quoted against.
 * Runs `iters` dependent-chain FMAs x 8 chains per thread; flops = grid*block*iters*8*2. */
int mpk_fma_peak(int dtype, int blocks, int threads, int64_t iters, double *sink_dev, void *stream);

/* Store-bandwidth micro-benchmark: `blocks` x 256 threads write `bytes` (a multiple of 16) to
 * dst_dev with 16-byte vector stores, grid-strided.  mode 0 st.global, 1 st.global.cs (what the
 * row-writing kernels use), 2 .cg, 3 .wt.  bench.py / scripts/microbench.py time it with CUDA
 * events to get the write-only HBM ceiling the trajectory kernel's roofline is quoted against. */
int mpk_store_peak(void *dst_dev, int64_t bytes, int mode, int blocks, void *stream);

/* ---- collision / limit post-processing hook of joint_trajectory -------------------
 * Replaces the per-row host loop of planning/collision_host.py:40-88 -- URDF.link_fk over the link tree
 * (urdf/core.py:532-633), CollisionChecker.check_collision (potential_field/collision.py:162-221: boxes
 * of the transformed hull points, pairwise overlap outside the allowed-collision set) and the attractive
 * potential-field step (potential_field/fields.py:112-170 with no obstacles).
 *
 * The collision model is packed once on the host and uploaded by the caller (no hidden allocation):
 *   link_joint (L) int32      actuated joint each link hangs on, -1 = fixed to the base
 *   link_home  (L, 4, 4)      link poses at the zero configuration (URDF.link_fk(zeros))
 *   acm        (L, L) uint8   1 = pair excluded (potential_field/adjacency.py:9-30)
 *   hull_link  (H) int32, hull_count (H) int32, hull_points (sum counts, 3) float64
 *                             ConvexHull.points of every link that has geometry, link frame, in the
 *                             checker's insertion order
 * mpk_collision_model_bytes gives the size of the packed model, mpk_collision_model_pack fills `out`
 * (host memory).  Launchers take the packed model twice: `model_host` (only its header is read, for
 * validation) and `model_dev` (the same bytes in device memory). */
size_t mpk_collision_model_bytes(int n, int L, int H, int64_t npoints);
int mpk_collision_model_pack(const mpk_robot *rb, int L, const int32_t *link_joint, const double *link_home,
                             const uint8_t *acm, int H, const int32_t *hull_link, const int32_t *hull_count,
                             const double *hull_points, void *out, size_t out_bytes);
/* URDF.link_fk_batch (urdf/core.py:577-633): T dev (P, L, 4, 4) float64, theta dev (P, n) theta_dtype */
int mpk_link_fk_batch(const mpk_robot *rb, const void *model_host, const void *model_dev, int64_t P,
                      const void *theta, int theta_dtype, double *T, void *stream);
/* CollisionChecker.check_collision per configuration: flags dev (P) uint8 */
int mpk_self_collision_aabb(const mpk_robot *rb, const void *model_host, const void *model_dev, int64_t P,
                            const void *theta, int theta_dtype, uint8_t *flags, void *stream);
/* _apply_collision_avoidance_cpu (collision_host.py:40-88): rows dev (P, n) float32 are nudged IN PLACE,
 * row p towards goal[p / rows_per_goal] (goal dev (G, n) float32: thetaend of the row's trajectory):
 * while the row collides and fewer than max_iterations steps were taken,
 *   row <- row - step * (attractive_gain * (row - goal))       (float32, every operation rounded)
 * iterations dev (P) int32 (steps taken) and flags dev (P) uint8 (still colliding) may be NULL.
 * The reference uses step = 0.01, max_iterations = 100, attractive_gain = 1. */
int mpk_collision_avoidance(const mpk_robot *rb, const void *model_host, const void *model_dev, int64_t P,
                            float *rows, const float *goal, int64_t rows_per_goal, double attractive_gain,
                            double step, int max_iterations, int32_t *iterations, uint8_t *flags, void *stream);

/* The reference's LEGACY dynamics path: a ManipulatorDynamics built without Mlist_per_link
 * falls back to _mass_matrix_legacy (dynamics/mass_matrix.py:101-132) and _gravity_forces_legacy
 * (dynamics/forces.py:135-154), with finite-difference Coriolis forces (dynamics/cache.py:23-56), inverse and
 * forward dynamics (dynamics/id_fd.py:16-83) on top.  Documented by the reference as incorrect physics;
 * reproduced for parity, not tuned.  S_list (6, n), M (4, 4), Glist (n, 6, 6) host float64;
 *   mode 0 mass matrix  out (P, n, n)          mode 1 gravity forces  out (P, n)
 *   mode 2 velocity-quadratic forces (dtheta)   mode 3 inverse dynamics (dtheta, third = ddtheta)
 *   mode 4 forward dynamics (dtheta, third = tau);   theta / dtheta / third / out dev float64;
 *   g host (3); Ftip host (6) or NULL; Ftip_rows dev (P, 6) or NULL (overrides Ftip) */
int mpk_legacy_dynamics(int n, const double *S_list, const double *M, const double *Glist, int mode, int64_t P,
                        const double *theta, const double *dtheta, const double *third, const double *g,
                        const double *Ftip, const double *Ftip_rows, double *out, void *stream);

/* Result buffers shared by the processes of one box (one process per GPU), SURVEY.md 8e: the
 * rank that collects a sharded result allocates it with mpk_peer_alloc and hands the 64-byte handle
 * to the other ranks (any channel: torch.distributed's object broadcast in the Python host); they
 * map it with mpk_peer_open -- the driver enables peer access over NVLink -- and pass the mapped
 * pointer (plus their row offset) as the OUTPUT pointer of the launchers above.  The kernels' row
 * stores then go straight into the collecting GPU's HBM: no gather step after the kernel.
 * All four act on the CURRENT CUDA device. */
#define MPK_PEER_HANDLE_BYTES 64
int mpk_peer_alloc(size_t bytes, void **dev_ptr, unsigned char *handle /* [MPK_PEER_HANDLE_BYTES] out */);
int mpk_peer_free(void *dev_ptr);
int mpk_peer_open(const unsigned char *handle /* [MPK_PEER_HANDLE_BYTES] */, void **dev_ptr);
int mpk_peer_close(void *dev_ptr);

#ifdef __cplusplus
}
#endif
#endif /* MPK_H */
