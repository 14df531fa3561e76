"""Parity of the CUDA path (through torch.ops.mpk -> C ABI -> sm_100a kernels) with the oracle.

Tolerances (BASELINE.json north_star):
  * fp64 kernels: per-vector inf-norm relative 1e-9, i.e. max|d| <= 1e-9 * max(1, |ref|_inf)
    (element-wise relative is unattainable even reference-vs-itself);
    against the reference's own goldens its own tolerances rtol 1e-7 / atol 1e-9, 1e-8;
  * float32 trajectory-level outputs: within one float32 ulp-scale of the oracle (rtol 3e-7);
  * time-scaling / trajectory rows: bit-exact.
"""

import numpy as np
import pytest

from conftest import load_golden, load_pack, planar_2r_pack, random_general_pack

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROBOTS = ["ur5", "panda", "iiwa14", "xarm6"]


def _rel_rows(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    sc = np.maximum(1.0, np.abs(ref).reshape(ref.shape[0], -1).max(1))
    d = np.abs(np.asarray(got, dtype=np.float64) - ref).reshape(ref.shape[0], -1).max(1)
    return float((d / sc).max())


def _bits_equal(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.fixture(scope="module")
def robots():
    from manipulapy_b200 import load_robot

    assert torch.cuda.is_available()
    return {name: load_robot(name) for name in ROBOTS}


# ---------------------------------------------------------------------------------------------
# reference golden vectors
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("robot", ["ur5", "panda", "iiwa14"])
def test_dynamics_golden(robots, robot):
    g = load_golden(f"dynamics_{robot}")
    dyn = robots[robot].dynamics
    th, dth, ddth = g["thetas"], g["dthetas"], g["ddthetas"]
    np.testing.assert_allclose(dyn.forward_kinematics(th), g["forward_kinematics"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(dyn.jacobian(th), g["jacobian"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(dyn.mass_matrix(th), g["mass_matrix"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(dyn.gravity_forces(th, g["g"]), g["gravity_forces"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(dyn.velocity_quadratic_forces(th, dth), g["velocity_quadratic_forces"],
                               rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(dyn.inverse_dynamics(th, dth, ddth, g["g"], g["ftips"]), g["inverse_dynamics"],
                               rtol=1e-7, atol=1e-8)
    i = g["fd_index"]
    dd = dyn.forward_dynamics(th[i], dth[i], g["fd_tau"], g["g"], g["ftips"][i])
    assert _rel_rows(dd, g["forward_dynamics"]) < 1e-9
    # single-sample calls keep the reference's shapes and dtype
    M1 = dyn.mass_matrix(th[3])
    n = th.shape[1]
    assert M1.shape == (n, n) and M1.dtype == np.float64
    assert np.array_equal(M1, dyn.mass_matrix(th[3:4])[0])
    assert dyn.forward_kinematics(th[3]).shape == (4, 4) and dyn.jacobian(th[3]).shape == (6, n)
    tau1 = dyn.inverse_dynamics(th[5], dth[5], ddth[5], g["g"], g["ftips"][5])
    np.testing.assert_allclose(tau1, g["inverse_dynamics"][5], rtol=1e-7, atol=1e-8)


def test_joint_trajectory_golden(robots):
    g = load_golden("trajectory")
    rb = robots["ur5"]
    assert np.array_equal(rb.joint_limits, g["joint_limits"])
    planner = rb.planner()
    for name in ("cfg1", "cubic50", "two", "clipped", "odd_tf"):
        Tf, N, method = g[f"{name}_args"]
        r = planner.joint_trajectory(g[f"{name}_start"], g[f"{name}_end"], Tf, int(N), int(method))
        for k in ("positions", "velocities", "accelerations"):
            got, ref = r[k], g[f"{name}_{k}"]
            assert got.dtype == np.float32 and got.shape == ref.shape
            neq = got.view(np.uint32) != ref.view(np.uint32)
            # the reference's Numba fastmath kernel leaves O(1e-16) residues at exact-cancellation
            # points (tau = 0.5, 1) where IEEE arithmetic gives 0: see tests/test_oracle_golden.py
            assert np.all(got[neq] == 0) and np.all(np.abs(ref[neq]) <= 1e-12), (name, k)
        if name in ("cfg1", "odd_tf"):
            assert all(_bits_equal(r[k], g[f"{name}_{k}"]) for k in r)
    Tf, N, method = g["batch_args"]
    r = planner.batch_joint_trajectory(g["batch_start"], g["batch_end"], Tf, int(N), int(method))
    for k in ("positions", "velocities", "accelerations"):
        neq = r[k].view(np.uint32) != g[f"batch_{k}"].view(np.uint32)
        assert np.all(r[k][neq] == 0) and np.all(np.abs(g[f"batch_{k}"][neq]) <= 1e-12)


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_inverse_dynamics_trajectory_golden(robots, robot):
    g = load_golden("id_trajectory")
    rb = robots[robot]
    tl = g[f"{robot}_torque_limits"]
    planner = rb.planner(torque_limits=tl)
    th, dth, ddth = g[f"{robot}_theta"], g[f"{robot}_dtheta"], g[f"{robot}_ddtheta"]
    t0 = planner.inverse_dynamics_trajectory(th, dth, ddth)
    t1 = planner.inverse_dynamics_trajectory(th, dth, ddth, g[f"{robot}_g1"], g[f"{robot}_ftip1"])
    lo32, hi32 = tl[:, 0].astype(np.float32), tl[:, 1].astype(np.float32)
    for got, ref in ((t0, g[f"{robot}_tau_default"]), (t1, g[f"{robot}_tau_g1_ftip1"])):
        assert got.dtype == np.float32 and got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=3e-7, atol=1e-7)
        assert np.array_equal(got == lo32, ref == lo32) and np.array_equal(got == hi32, ref == hi32)


@pytest.mark.parametrize("robot", ["iiwa14", "ur5"])
def test_forward_dynamics_trajectory_golden(robots, robot):
    g = load_golden("fd_trajectory")
    rb = robots[robot]
    planner = rb.planner()
    assert np.array_equal(rb.joint_limits, g[f"{robot}_joint_limits"])
    dt, intres = g[f"{robot}_a_args"]
    r = planner.forward_dynamics_trajectory(g[f"{robot}_a_theta0"], g[f"{robot}_a_dtheta0"], g[f"{robot}_a_tau"],
                                            [0, 0, -9.81], None, dt, int(intres))
    for k in ("positions", "velocities", "accelerations"):
        assert r[k].dtype == np.float32 and _rel_rows(r[k], g[f"{robot}_a_{k}"]) <= 1e-6, k
    dt, intres = g[f"{robot}_b_args"]
    r = planner.forward_dynamics_trajectory(g[f"{robot}_b_theta0"], g[f"{robot}_b_dtheta0"], g[f"{robot}_b_tau"],
                                            g[f"{robot}_b_g"], g[f"{robot}_b_ftip"], dt, int(intres))
    for k in ("positions", "velocities", "accelerations"):
        assert _rel_rows(r[k], g[f"{robot}_b_{k}"]) <= 1e-6, k
    hi32 = rb.joint_limits[:, 1].astype(np.float32)
    assert np.array_equal(r["positions"] == hi32, g[f"{robot}_b_positions"] == hi32)
    with pytest.raises(IndexError):
        planner.forward_dynamics_trajectory(np.zeros(rb.num_joints), np.zeros(rb.num_joints),
                                            np.zeros((0, rb.num_joints)), [0, 0, -9.81], None, 1e-3, 1)


# ---------------------------------------------------------------------------------------------
# seeded random inputs against the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("robot", ROBOTS)
def test_kinematics_dynamics_vs_oracle(robots, oracle_factory, robot):
    rb, o = robots[robot], oracle_factory(robot)
    n = rb.num_joints
    rng = np.random.default_rng(21)
    P = 4099  # ragged: not a multiple of the 128-thread block
    th = rng.uniform(rb.joint_limits[:, 0], rb.joint_limits[:, 1], (P, n))
    dth, ddth = rng.uniform(-2, 2, (P, n)), rng.uniform(-5, 5, (P, n))
    g, ft = np.array([0.3, -0.2, -9.81]), rng.uniform(-10, 10, 6)
    dyn = rb.dynamics
    T, J = dyn.forward_kinematics_and_jacobian(th)
    assert np.abs(T - o.forward_kinematics(th)).max() < 1e-12
    assert np.abs(J - o.jacobian(th)).max() < 1e-12
    assert np.array_equal(T, dyn.forward_kinematics(th)) and np.array_equal(J, dyn.jacobian(th))
    assert _rel_rows(dyn.inverse_dynamics(th, dth, ddth, g, ft), o.inverse_dynamics(th, dth, ddth, g, ft, analytic=True)) < 1e-9
    rows = rng.uniform(-10, 10, (P, 6))
    assert _rel_rows(dyn.inverse_dynamics(th, dth, ddth, g, rows), o.inverse_dynamics(th, dth, ddth, g, rows, analytic=True)) < 1e-9
    assert _rel_rows(dyn.gravity_forces(th, g), o.gravity_forces(th, g, analytic=True)) < 1e-9
    assert _rel_rows(dyn.velocity_quadratic_forces(th, dth), o.velocity_quadratic_forces(th, dth, analytic=True)) < 1e-9
    assert _rel_rows(dyn.mass_matrix(th[:513]), o.mass_matrix(th[:513])) < 1e-9
    tau = rng.uniform(-20, 20, (513, n))
    assert _rel_rows(dyn.forward_dynamics(th[:513], dth[:513], tau, g, ft),
                     o.forward_dynamics(th[:513], dth[:513], tau, g, ft, analytic=True)) < 1e-9
    # a small sample against the LITERAL reference algorithm (finite-difference Coriolis noise ~1e-9 abs)
    lit = o.inverse_dynamics(th[:16], dth[:16], ddth[:16], g, ft, analytic=False)
    np.testing.assert_allclose(dyn.inverse_dynamics(th[:16], dth[:16], ddth[:16], g, ft), lit, rtol=1e-7, atol=1e-7)
    # float32 theta storage is upcast exactly
    th32 = th.astype(np.float32)
    assert np.array_equal(dyn.mass_matrix(th32[:64]), dyn.mass_matrix(th32[:64].astype(np.float64)))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8])
def test_general_inertia_robot_vs_oracle(n):
    """Arbitrary unit screws, one non-unit prismatic joint, full symmetric 6x6 inertias."""
    from manipulapy_b200 import ManipulatorDynamics
    from oracle import Oracle

    p = random_general_pack(n, seed=100 + n)
    o = Oracle(p["S_list"], p["M"], p["Glist"], p["Mlist_per_link"])
    dyn = ManipulatorDynamics(p["M"], None, None, None, p["S_list"], None, p["Glist"], p["Mlist_per_link"])
    assert not dyn.robot.rigid
    rng = np.random.default_rng(n)
    P = 300
    th, dth, ddth = rng.uniform(-2, 2, (P, n)), rng.uniform(-2, 2, (P, n)), rng.uniform(-3, 3, (P, n))
    g, ft = np.array([1.0, 2.0, -9.0]), rng.uniform(-5, 5, 6)
    assert np.abs(dyn.forward_kinematics(th) - o.forward_kinematics(th)).max() < 1e-12
    assert np.abs(dyn.jacobian(th) - o.jacobian(th)).max() < 1e-12
    assert _rel_rows(dyn.inverse_dynamics(th, dth, ddth, g, ft), o.inverse_dynamics(th, dth, ddth, g, ft, analytic=True)) < 1e-9
    assert _rel_rows(dyn.mass_matrix(th), o.mass_matrix(th)) < 1e-9
    assert _rel_rows(dyn.gravity_forces(th, g), o.gravity_forces(th, g)) < 1e-9
    tau = rng.uniform(-5, 5, (P, n))
    assert _rel_rows(dyn.forward_dynamics(th, dth, tau, g, ft), o.forward_dynamics(th, dth, tau, g, ft, analytic=True)) < 1e-9


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_rigid_and_general_kernels_agree(robots, robot):
    from manipulapy_b200 import ManipulatorDynamics

    p = load_pack(robot)
    gen = ManipulatorDynamics(p["M"], None, None, None, p["S_list"], None, p["Glist"], p["Mlist_per_link"],
                              force_general_inertia=True)
    rig = robots[robot].dynamics
    assert rig.robot.rigid and not gen.robot.rigid
    rng = np.random.default_rng(5)
    n = rig.num_joints
    th, dth, ddth = rng.uniform(-3, 3, (777, n)), rng.uniform(-2, 2, (777, n)), rng.uniform(-5, 5, (777, n))
    assert _rel_rows(gen.inverse_dynamics(th, dth, ddth, [0, 0, -9.81]), rig.inverse_dynamics(th, dth, ddth, [0, 0, -9.81])) < 1e-11
    assert _rel_rows(gen.mass_matrix(th), rig.mass_matrix(th)) < 1e-11


def test_planar_2r_known_answers():
    from manipulapy_b200 import ManipulatorDynamics

    p = planar_2r_pack()
    dyn = ManipulatorDynamics(p["M"], None, None, None, p["S_list"], None, p["Glist"], p["Mlist_per_link"])
    th = np.array([0.0, np.pi / 2])
    c2 = np.cos(th[1])
    np.testing.assert_allclose(dyn.mass_matrix(th), [[1 + (2 + 2 * c2), 1 + c2], [1 + c2, 1.0]], atol=1e-12)
    np.testing.assert_allclose(dyn.gravity_forces(th, [-9.81, 0, 0]), [-9.81, -9.81], atol=1e-12)
    np.testing.assert_allclose(dyn.gravity_forces(th, [0, -9.81, 0]), [19.62, 0.0], atol=1e-12)
    np.testing.assert_allclose(dyn.gravity_forces(th, [0, 0, -9.81]), [0.0, 0.0], atol=1e-12)
    with pytest.raises(ValueError):
        dyn.forward_kinematics(th, frame="nope")


@pytest.mark.parametrize("robot,B,N,method", [("ur5", 37, 301, 5), ("iiwa14", 5, 129, 3), ("panda", 3, 128, 5),
                                              ("ur5", 300, 2, 5), ("ur5", 4, 1, 5), ("xarm6", 2, 50, 7)])
def test_batch_trajectory_bit_exact_vs_oracle(robots, robot, B, N, method):
    from oracle import Oracle

    rb = robots[robot]
    n = rb.num_joints
    rng = np.random.default_rng(B * 1000 + N)
    s, e = rng.uniform(-4, 4, (B, n)), rng.uniform(-4, 4, (B, n))
    planner = rb.planner()
    for dt in (np.float64, np.float32):
        r = planner.batch_joint_trajectory(s.astype(dt), e.astype(dt), 1.7, N, method)
        ref = Oracle.joint_trajectory(s.astype(dt), e.astype(dt), 1.7, N, method, rb.joint_limits)
        for k in ref:
            a, b = r[k], ref[k]
            assert a.shape == (B, N, n) and a.dtype == np.float32
            if N == 1:  # 0 * inf = NaN on the reference's CPU path
                assert np.array_equal(np.isnan(a), np.isnan(b))
            else:
                assert _bits_equal(a, b), (k, dt)
    r = planner.batch_joint_trajectory(np.zeros((0, n)), np.zeros((0, n)), 1.0, 10, 5)
    assert r["positions"].shape == (0, 10, n)


@pytest.mark.parametrize("robot", ["ur5", "iiwa14", "panda"])
def test_fused_trajectory_inverse_dynamics_equals_two_calls(robots, oracle_factory, robot):
    rb, o = robots[robot], oracle_factory(robot)
    n = rb.num_joints
    rng = np.random.default_rng(8)
    B, N = 19, 333
    s, e = rng.uniform(-3, 3, (B, n)), rng.uniform(-3, 3, (B, n))
    tl = np.array([[-60.0, 55.0]] * n)
    planner = rb.planner(torque_limits=tl)
    ft = [1.0, -2.0, 0.5, 3.0, 0.0, -1.0]
    tr = planner.batch_joint_trajectory(s, e, 2.0, N, 5)
    two = planner.inverse_dynamics_trajectory(tr["positions"], tr["velocities"], tr["accelerations"], [0, 0, -9.81], ft)
    tau, tr2 = planner.trajectory_inverse_dynamics(s, e, 2.0, N, 5, [0, 0, -9.81], ft, return_trajectory=True)
    assert two.shape == (B, N, n) and _bits_equal(tau, two)
    assert all(_bits_equal(tr[k], tr2[k]) for k in tr)
    assert _bits_equal(planner.trajectory_inverse_dynamics(s, e, 2.0, N, 5, [0, 0, -9.81], ft), two)
    ref = o.inverse_dynamics_trajectory(tr["positions"].reshape(-1, n), tr["velocities"].reshape(-1, n),
                                        tr["accelerations"].reshape(-1, n), [0, 0, -9.81], ft, tl, analytic=True)
    np.testing.assert_allclose(two.reshape(-1, n), ref, rtol=3e-7, atol=1e-7)
    assert (two == np.float32(55.0)).any() or (two == np.float32(-60.0)).any()


@pytest.mark.parametrize("robot,B,N,intres", [("iiwa14", 70, 60, 1), ("ur5", 33, 40, 3), ("panda", 9, 30, 2)])
def test_batched_rollouts_vs_oracle(robots, oracle_factory, robot, B, N, intres):
    rb, o = robots[robot], oracle_factory(robot)
    n = rb.num_joints
    rng = np.random.default_rng(B)
    lo, hi = rb.joint_limits[:, 0], rb.joint_limits[:, 1]
    th0 = rng.uniform(0.5 * lo, 0.5 * hi, (B, n))
    th0[0] = hi - 1e-4  # start next to the upper limit so the clip is exercised
    dth0 = rng.uniform(-0.5, 0.5, (B, n))
    dth0[0] = 2.0
    tau = rng.uniform(-20, 20, (B, N, n))
    ftip = rng.uniform(-3, 3, (B, N, 6))
    planner = rb.planner()
    for fm in (None, ftip):
        r = planner.forward_dynamics_trajectory(th0, dth0, tau, [0, 0, -9.81], fm, 1e-3, intres)
        ref = o.forward_dynamics_trajectory(th0, dth0, tau, [0, 0, -9.81], fm, 1e-3, intres, rb.joint_limits, analytic=True)
        for k in ref:
            assert r[k].shape == (B, N, n) and r[k].dtype == np.float32
            assert _rel_rows(r[k].reshape(B * N, n), ref[k].reshape(B * N, n)) <= 1e-6, k
        hi32 = hi.astype(np.float32)
        assert (ref["positions"][0, 1:] == hi32).any()
        assert np.array_equal(r["positions"] == hi32, ref["positions"] == hi32)
    # float32 torque storage is upcast exactly
    r32 = planner.forward_dynamics_trajectory(th0, dth0, tau.astype(np.float32), [0, 0, -9.81], None, 1e-3, intres)
    r64 = planner.forward_dynamics_trajectory(th0, dth0, tau.astype(np.float32).astype(np.float64), [0, 0, -9.81], None, 1e-3, intres)
    assert all(_bits_equal(r32[k], r64[k]) for k in r32)


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_rollout_kernels_agree(robots, oracle_factory, robot):
    """A batch that fits the GPU in one wave runs each Euler step split across warps -- three per 32
    rollouts up to 64 x SMs rollouts (fd_rollout_trio_kernel, one or two groups per block), two up to
    128 x SMs (fd_rollout_pair_kernel); up to 256 x SMs one warp per 32 rollouts (fd_rollout_kernel), beyond the pair
    kernel again (compiled for 6 blocks per SM).
    The same rollouts through all of them -- one call of 24,000 against calls of 4,000, 6,000 and 14,000 --
    give the same bits (so a batch sharded over GPUs equals the batch on one GPU whichever kernel each
    shard takes), ragged last block, intRes = 2, and all agree with the oracle on sampled rollouts."""
    rb, o = robots[robot], oracle_factory(robot)
    n = rb.num_joints
    B, N = 24000, 40
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    assert B > 4 * 32 * sms >= 14000 > 2 * 32 * sms >= 6000 > 32 * sms >= 4000
    rng = np.random.default_rng(77)
    lo, hi = rb.joint_limits[:, 0], rb.joint_limits[:, 1]
    th0 = rng.uniform(0.5 * lo, 0.5 * hi, (B, n))
    dth0 = rng.uniform(-0.5, 0.5, (B, n))
    tau = rng.uniform(-10, 10, (B, N, n)).astype(np.float32)
    planner = rb.planner()
    big = planner.forward_dynamics_trajectory(th0, dth0, tau, [0, 0, -9.81], None, 1e-3, 2)
    cuts = [0, 4000, 10000, 24000]
    parts = [planner.forward_dynamics_trajectory(th0[i:j], dth0[i:j], tau[i:j], [0, 0, -9.81], None, 1e-3, 2)
             for i, j in zip(cuts[:-1], cuts[1:])]
    for k in big:
        small = np.concatenate([p[k] for p in parts])
        assert _bits_equal(small, big[k]), k
    # beyond one wave of the single-warp kernel (256 x SMs rollouts) the pair kernel compiled for 6 blocks per SM runs:
    # 40,000 rollouts (two copies of the first 20,000) against the 20,000 through the single-warp kernel
    assert 2 * 20000 > 8 * 32 * sms >= 20000 > 4 * 32 * sms
    half = planner.forward_dynamics_trajectory(th0[:20000], dth0[:20000], tau[:20000, :12], [0, 0, -9.81], None, 1e-3, 2)
    twice = planner.forward_dynamics_trajectory(np.tile(th0[:20000], (2, 1)), np.tile(dth0[:20000], (2, 1)),
                                                np.tile(tau[:20000, :12], (2, 1, 1)), [0, 0, -9.81], None, 1e-3, 2)
    for k in half:
        assert _bits_equal(twice[k][:20000], half[k]) and _bits_equal(twice[k][20000:], half[k]), k
    odd = planner.forward_dynamics_trajectory(th0[:1001], dth0[:1001], tau[:1001], [0, 0, -9.81], None, 1e-3, 2)
    idx = np.array([0, 31, 32, 999, 1000])
    ref = o.forward_dynamics_trajectory(th0[idx], dth0[idx], tau[idx].astype(np.float64), [0, 0, -9.81], None, 1e-3, 2,
                                        rb.joint_limits, analytic=True)
    for k in ref:
        assert _rel_rows(odd[k][idx].reshape(-1, n), ref[k].reshape(-1, n)) <= 1e-6, k
        assert _rel_rows(big[k][idx].reshape(-1, n), ref[k].reshape(-1, n)) <= 1e-6, k


def test_rollout_kernels_long_horizon(robots):
    """The three rollout kernels over 1000 steps (the bench's torque distribution): the float32 positions
    agree bit for bit; velocities / accelerations may differ in the LAST float32 bit in a handful of entries
    (measured: 1-4 of 14.3 M): the kernels' float64 states drift apart by a few ulps over hundreds of steps --
    the same expressions, compiled in different kernels, are not contracted to FMAs identically -- and a
    float64 ulp occasionally decides a float32 rounding."""
    rb = robots["iiwa14"]
    n, B0, N = rb.num_joints, 1024, 1000
    rng = np.random.default_rng(5)
    lo, hi = rb.joint_limits[:, 0], rb.joint_limits[:, 1]
    th0 = rng.uniform(0.5 * lo, 0.5 * hi, (B0, n))
    dth0 = rng.uniform(-0.5, 0.5, (B0, n))
    amp = np.array([4.0, 4.0, 2.0, 2.0, 0.4, 0.2, 0.08])
    dyn = rb.dynamics
    tau = (np.asarray(dyn.gravity_forces(th0))[:, None, :] + rng.uniform(-0.5, 0.5, (B0, N, n)) * amp).astype(np.float32)
    planner = rb.planner()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    outs = []
    for reps in (1, 2 * 32 * sms // B0 + 1, 4 * 32 * sms // B0 + 1):  # three warps, two warps, one warp per 32 rollouts
        r = planner.forward_dynamics_trajectory(np.tile(th0, (reps, 1)), np.tile(dth0, (reps, 1)), np.tile(tau, (reps, 1, 1)),
                                                [0, 0, -9.81], None, 1e-3, 1)
        outs.append({k: np.ascontiguousarray(v[:B0]) for k, v in r.items()})
    for other in outs[1:]:
        for k in outs[0]:
            a, b = outs[0][k].view(np.int32).astype(np.int64), other[k].view(np.int32).astype(np.int64)
            d = np.abs(a - b)
            assert d.max() <= 1 and (d != 0).mean() <= 1e-5, (k, int(d.max()), float((d != 0).mean()))


def test_device_resident_path(robots):
    rb = robots["ur5"]
    planner = rb.planner()
    s = torch.rand(6, 6, dtype=torch.float64, device="cuda")
    e = torch.rand(6, 6, dtype=torch.float64, device="cuda")
    tr = planner.batch_joint_trajectory(s, e, 2.0, 100, 5)
    assert all(isinstance(v, torch.Tensor) and v.is_cuda and v.dtype == torch.float32 for v in tr.values())
    tau = planner.inverse_dynamics_trajectory(tr["positions"], tr["velocities"], tr["accelerations"])
    assert tau.is_cuda and tau.shape == (6, 100, 6)
    host = planner.batch_joint_trajectory(s.cpu().numpy(), e.cpu().numpy(), 2.0, 100, 5)
    assert _bits_equal(tr["positions"].cpu().numpy(), host["positions"])
    M = rb.dynamics.mass_matrix(tr["positions"].reshape(-1, 6)[:10])
    assert M.is_cuda and M.shape == (10, 6, 6)
    st = planner.get_performance_stats()
    assert st["gpu_calls"] >= 3 and st["cpu_calls"] == 0


def test_body_frame_kinematics(robots):
    """frame="body" FK / Jacobian against the reference goldens, against the numpy restatement
    on a batch, and J_b = Ad(T^-1) J_s for the URDF robots (whose B_list is consistent)."""
    from manipulapy_b200 import SerialManipulator
    from manipulapy_b200.kinematics import _adjoint
    from oracle import Oracle

    g = load_golden("body_kinematics")
    for k in ("ur5", "free"):
        sm = SerialManipulator(M_list=g[f"{k}_M"], S_list=g[f"{k}_S"], B_list=g[f"{k}_B"])
        th = g[f"{k}_theta"]
        np.testing.assert_allclose(sm.forward_kinematics(th, frame="body"), g[f"{k}_T"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(sm.jacobian(th, frame="body"), g[f"{k}_J"], rtol=0, atol=1e-12)
        assert sm.forward_kinematics(th[0], "body").shape == (4, 4) and sm.jacobian(th[0], "body").shape == (6, th.shape[1])
        rng = np.random.default_rng(0)
        big = rng.uniform(-3, 3, (1025, th.shape[1]))
        T, J = sm.forward_kinematics_and_jacobian(big, frame="body")
        np.testing.assert_allclose(T[:200], Oracle.body_forward_kinematics(g[f"{k}_M"], g[f"{k}_B"], big[:200]), rtol=0, atol=1e-12)
        np.testing.assert_allclose(J[:200], Oracle.body_jacobian(g[f"{k}_B"], big[:200]), rtol=0, atol=1e-12)
    rb = robots["iiwa14"]
    th = np.random.default_rng(1).uniform(-2, 2, (300, 7))
    Ts, Js = rb.dynamics.forward_kinematics_and_jacobian(th)
    Tb, Jb = rb.dynamics.forward_kinematics_and_jacobian(th, frame="body")
    np.testing.assert_allclose(Tb, Ts, rtol=0, atol=1e-12)
    ref = np.stack([_adjoint(np.linalg.inv(T)) @ J for T, J in zip(Ts, Js)])
    np.testing.assert_allclose(Jb, ref, rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        rb.dynamics.jacobian(th[0], frame="tool")


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_batched_inverse_kinematics(robot):
    """iterative_inverse_kinematics: the reference's goldens (success, iteration counts, solutions),
    then 20,000 targets in one launch: every converged solution reproduces its target pose."""
    from manipulapy_b200 import SerialManipulator

    g = load_golden("inverse_kinematics")
    lim = [tuple(r) for r in g[f"{robot}_limits"]]
    sm = SerialManipulator(M_list=g[f"{robot}_M"], S_list=g[f"{robot}_S"], joint_limits=lim)
    n = sm.num_joints
    for i, (Td, seed, par) in enumerate(zip(g[f"{robot}_T"], g[f"{robot}_seed"], g[f"{robot}_params"])):
        # single-target calls take the stagnation-restart noise from NumPy's global generator like the
        # reference (seeded as the golden run was), so every run is reproduced, restarts included
        np.random.seed(100 + i)
        th, ok, it = sm.iterative_inverse_kinematics(Td, seed, max_iterations=int(par[0]), damping=par[1],
                                                     step_cap=par[2], weight_orientation=par[3], weight_position=par[4])
        assert th.shape == (n,) and isinstance(ok, bool) and isinstance(it, int)
        assert ok == bool(g[f"{robot}_success"][i]) and it == int(g[f"{robot}_iterations"][i]), i
        np.testing.assert_allclose(th, g[f"{robot}_theta"][i], rtol=0, atol=1e-7)
    rng = np.random.default_rng(2)
    P = 20001
    lo, hi = g[f"{robot}_limits"][:, 0], g[f"{robot}_limits"][:, 1]
    tgt = rng.uniform(0.5 * lo, 0.5 * hi, (P, n))
    Td = sm.forward_kinematics(tgt)
    th, ok, it = sm.iterative_inverse_kinematics(Td, tgt + rng.uniform(-0.3, 0.3, (P, n)), max_iterations=400)
    assert th.shape == (P, n) and ok.dtype == bool and it.dtype == np.int32
    assert ok.mean() > 0.95  # (the method itself stalls on a few per cent of random 7-DOF targets)
    T = sm.forward_kinematics(th)
    assert np.abs(T[ok] - Td[ok])[:, :3, 3].max() < 2e-6 and np.abs(T[ok] - Td[ok])[:, :3, :3].max() < 2e-6
    assert (it[ok] <= 400).all() and (it[~ok] == 401).all()
    assert (th >= lo - 1e-12).all() and (th <= hi + 1e-12).all()
    # the re-packing ladder (stragglers re-packed after 16, 32, 64, ... iterations) changes no iterate
    from manipulapy_b200 import _native
    dev = torch.device("cuda")
    args = (sm.robot.handle, torch.from_numpy(Td).to(dev), torch.from_numpy(tgt + 0.25).to(dev), 1e-6, 1e-6, 400, 2e-2,
            0.3, 1.0, 1.0, torch.from_numpy(np.ascontiguousarray(g[f"{robot}_limits"], dtype=np.float64)), 7)
    one = _native.ops().inverse_kinematics_dls(*args, False)
    two = _native.ops().inverse_kinematics_dls(*args, True)
    assert all(bool(torch.equal(x, y)) for x, y in zip(one, two))
    assert 0 < int((two[2] > 64).sum()) < P  # some targets did go through the second phase
    with pytest.raises(NotImplementedError):
        sm.iterative_inverse_kinematics(Td[0], tgt[0], plot_residuals=True)


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_inverse_kinematics_modes_and_front_ends(robot):
    """adaptive_tuning / backtracking (kinematics/ik.py:215-229, 253-276) and the smart_ / robust_
    front ends (ik.py:327-598) against goldens generated from the unmodified reference: success,
    iteration counts, winning strategy and solutions for every run without a stagnation restart;
    then 5,000 random targets through each front end in batched launches."""
    from manipulapy_b200 import SerialManipulator, ik_helpers

    g = load_golden("inverse_kinematics_modes")
    lim = [tuple(r) for r in g[f"{robot}_limits"]]
    sm = SerialManipulator(M_list=g[f"{robot}_M"], S_list=g[f"{robot}_S"], joint_limits=lim)
    for mode in ("adaptive", "backtracking", "both"):
        kw = dict(adaptive_tuning=mode in ("adaptive", "both"), backtracking=mode in ("backtracking", "both"))
        # all targets of the golden in ONE launch (ragged iteration budgets: the last one separately)
        Td, seed, budget = g[f"{robot}_T"], g[f"{robot}_seed"], g[f"{robot}_max_iterations"]
        th, ok, it = sm.iterative_inverse_kinematics(Td[:-1], seed[:-1], max_iterations=int(budget[0]), **kw)
        assert np.array_equal(ok, g[f"{robot}_{mode}_success"][:-1])
        assert np.array_equal(it[ok], g[f"{robot}_{mode}_iterations"][:-1][ok])
        np.testing.assert_allclose(th[ok], g[f"{robot}_{mode}_theta"][:-1][ok], rtol=0, atol=1e-7)
        for i in range(len(Td)):  # and one by one, as the reference is called (generator seeded like the golden run)
            np.random.seed(200 + i)
            th1, ok1, it1 = sm.iterative_inverse_kinematics(Td[i], seed[i], max_iterations=int(budget[i]), **kw)
            assert ok1 == bool(g[f"{robot}_{mode}_success"][i]) and it1 == int(g[f"{robot}_{mode}_iterations"][i]), (mode, i)
            np.testing.assert_allclose(th1, g[f"{robot}_{mode}_theta"][i], rtol=0, atol=1e-7)

    g = load_golden("inverse_kinematics_front_ends")
    lim = [tuple(r) for r in g[f"{robot}_limits"]]
    sm = SerialManipulator(M_list=g[f"{robot}_M"], S_list=g[f"{robot}_S"], joint_limits=lim)
    for i, Td in enumerate(g[f"{robot}_T"]):
        np.random.seed(300 + i)
        th, ok, it = sm.smart_inverse_kinematics(Td, max_iterations=120)
        assert isinstance(ok, bool) and isinstance(it, int) and th.shape == (len(lim),)
        # every case: fall-back starts and stagnation restarts draw from NumPy's generator like the reference
        assert ok == bool(g[f"{robot}_smart_success"][i]) and it == int(g[f"{robot}_smart_iterations"][i]), i
        np.testing.assert_allclose(th, g[f"{robot}_smart_theta"][i], rtol=0, atol=1e-7, err_msg=str(i))
        np.random.seed(400 + i)
        th, ok, it, win = sm.robust_inverse_kinematics(Td, max_attempts=4, max_iterations=120)
        assert ok == bool(g[f"{robot}_robust_success"][i]) and it == int(g[f"{robot}_robust_iterations"][i]), i
        assert win == str(g[f"{robot}_robust_strategy"][i])
        np.testing.assert_allclose(th, g[f"{robot}_robust_theta"][i], rtol=0, atol=1e-6, err_msg=str(i))
    with pytest.raises(ValueError):
        sm.smart_inverse_kinematics(g[f"{robot}_T"][0], strategy="nonsense")

    # batched: random reachable targets, no initial guess given
    rng = np.random.default_rng(5)
    lo, hi = g[f"{robot}_limits"][:, 0], g[f"{robot}_limits"][:, 1]
    P = 5000
    Td = sm.forward_kinematics(rng.uniform(0.8 * lo, 0.8 * hi, (P, len(lim))))
    np.random.seed(9)
    th, ok, it = sm.smart_inverse_kinematics(Td, max_iterations=150)
    assert th.shape == (P, len(lim)) and ok.mean() > 0.9
    assert ik_helpers.pose_error(sm.forward_kinematics(th[ok]), Td[ok]).max() < 1e-5
    plain = sm.smart_inverse_kinematics(Td, max_iterations=150, auto_fallback=False)[1]
    assert ok.sum() >= plain.sum()  # the fall-back starts only add solutions
    # the re-packing ladder carries the adaptive-tuning state of the stragglers through its queues
    from manipulapy_b200 import _native
    dev = torch.device("cuda")
    th0 = ik_helpers.workspace_heuristic_guess(Td, len(lim), lim)
    args = (sm.robot.handle, torch.from_numpy(Td).to(dev), torch.from_numpy(th0).to(dev), 1e-6, 1e-6, 300, 2e-2, 0.3,
            1.0, 1.0, torch.from_numpy(np.ascontiguousarray(g[f"{robot}_limits"], dtype=np.float64)), 3)
    one = _native.ops().inverse_kinematics_dls(*args, False, 3)
    two = _native.ops().inverse_kinematics_dls(*args, True, 3)
    assert all(bool(torch.equal(x, y)) for x, y in zip(one, two))
    assert 0 < int((two[2] > 64).sum()) < P
    np.random.seed(9)
    th, ok, it, win = sm.robust_inverse_kinematics(Td, max_attempts=6, max_iterations=150)
    assert ok.mean() > 0.95 and set(win) <= {"workspace_heuristic", "midpoint", "random", "none"}
    assert ik_helpers.pose_error(sm.forward_kinematics(th[ok]), Td[ok]).max() < 5e-3


def test_cartesian_trajectory(robots):
    """cartesian_trajectory against the reference goldens (float32 rounding), batched = per-pair,
    R R^T = 1 along 1,000,000 interpolated poses, exact end points."""
    from oracle import Oracle

    planner = robots["ur5"].planner()
    g = load_golden("cartesian_trajectory")
    names = ("generic5", "generic3", "method1", "same_R", "tiny", "near_pi", "pi_band", "pi_exact")
    for name in names:
        Tf, N, method = g[f"{name}_args"]
        got = planner.cartesian_trajectory(g[f"{name}_Xstart"], g[f"{name}_Xend"], float(Tf), int(N), int(method))
        for k in ("positions", "velocities", "accelerations", "orientations"):
            assert got[k].dtype == np.float32 and got[k].shape == g[f"{name}_{k}"].shape
            np.testing.assert_allclose(got[k], g[f"{name}_{k}"], rtol=3e-7, atol=1e-7, err_msg=f"{name} {k}")
    rng = np.random.default_rng(3)
    B, N = 1000, 1000
    def poses(B):
        Q, _ = np.linalg.qr(rng.normal(size=(B, 3, 3)))
        Q[:, :, 0] *= np.sign(np.linalg.det(Q))[:, None]
        T = np.tile(np.eye(4), (B, 1, 1))
        T[:, :3, :3] = Q
        T[:, :3, 3] = rng.uniform(-1, 1, (B, 3))
        return T
    Xs, Xe = poses(B), poses(B)
    r = planner.cartesian_trajectory(Xs, Xe, 2.0, N, 5)
    assert r["orientations"].shape == (B, N, 3, 3) and r["positions"].shape == (B, N, 3)
    R = r["orientations"].astype(np.float64)
    assert np.abs(np.einsum("bnij,bnkj->bnik", R, R) - np.eye(3)).max() < 1e-6
    np.testing.assert_allclose(R[:, 0], Xs[:, :3, :3], atol=1e-6)
    np.testing.assert_allclose(R[:, -1], Xe[:, :3, :3], atol=1e-6)
    np.testing.assert_allclose(r["positions"][:, -1], Xe[:, :3, 3], atol=1e-6)
    for b in (0, 499, 999):
        ref = Oracle.cartesian_trajectory(Xs[b], Xe[b], 2.0, N, 5)
        for k in ref:
            np.testing.assert_allclose(r[k][b], ref[k], rtol=3e-7, atol=1e-7)
    with pytest.raises(ZeroDivisionError):
        planner.cartesian_trajectory(Xs[0], Xe[0], 2.0, 1, 5)


@pytest.mark.parametrize("robot", ROBOTS)
def test_reference_property_tests(robots, robot):
    """The property tests of the reference's suite on batches: M symmetric positive definite and
    inverse dynamics at rest = gravity forces (tests/test_dynamics_golden.py:236-272); c(theta, 0) = 0
    and the inverse -> forward dynamics round trip (tests/test_dynamics.py:23-144, there atol 1e-3);
    FK(0) = M, R R^T = 1, det R = 1 (tests/test_kinematics.py:147-212); quintic trajectories start
    and end with zero velocity and acceleration (tests/test_v132_regressions.py:516-541)."""
    rb = robots[robot]
    dyn, n = rb.dynamics, rb.num_joints
    rng = np.random.default_rng(13)
    lo, hi = rb.joint_limits[:, 0], rb.joint_limits[:, 1]
    P = 3000
    th = rng.uniform(lo, hi, (P, n))
    dth, dd = rng.uniform(-2, 2, (P, n)), rng.uniform(-5, 5, (P, n))
    M = dyn.mass_matrix(th)
    assert np.array_equal(M, np.swapaxes(M, 1, 2))
    assert np.linalg.eigvalsh(M).min() > 0
    zero = np.zeros_like(th)
    grav = dyn.gravity_forces(th)
    # (gravity_forces runs the at-rest form of the recursion, the inverse dynamics the full one:
    # same value up to the rounding of terms that are multiplied by the zero velocities)
    rest = dyn.inverse_dynamics(th, zero, zero, [0, 0, -9.81], None)
    assert np.abs(rest - grav).max() <= 1e-13 * max(1.0, np.abs(grav).max())
    assert np.abs(dyn.velocity_quadratic_forces(th, zero)).max() == 0.0
    ft = rng.uniform(-5, 5, (P, 6))
    tau = dyn.inverse_dynamics(th, dth, dd, [0, 0, -9.81], ft)
    back = dyn.forward_dynamics(th, dth, tau, [0, 0, -9.81], ft)
    assert _rel_rows(back, dd) < 1e-9  # round trip through M^-1 (cond(M) up to ~3e4)
    T0 = dyn.forward_kinematics(np.zeros(n))
    np.testing.assert_allclose(T0, rb.M, rtol=0, atol=1e-13)
    T = dyn.forward_kinematics(th)
    R = T[:, :3, :3]
    assert np.abs(np.einsum("pij,pkj->pik", R, R) - np.eye(3)).max() < 1e-13
    assert np.abs(np.linalg.det(R) - 1).max() < 1e-13
    tr = rb.planner().batch_joint_trajectory(th[:50], rng.uniform(lo, hi, (50, n)), 2.0, 101, 5)
    for k in ("velocities", "accelerations"):
        assert np.abs(tr[k][:, 0]).max() == 0.0 and np.abs(tr[k][:, -1]).max() < 1e-4
    np.testing.assert_array_equal(tr["positions"][:, 0], th[:50].astype(np.float32))


def test_planner_utilities(robots):
    """calculate_derivatives (planning/trajectory_dynamics.py:710-731) and cleanup_gpu_memory."""
    planner = robots["ur5"].planner()
    x = np.random.default_rng(0).normal(size=(9, 6))
    v, a, j = planner.calculate_derivatives(x, 0.1)
    assert v.shape == (8, 6) and a.shape == (7, 6) and j.shape == (6, 6)
    assert np.array_equal(v, np.diff(x, axis=0) / 0.1) and np.array_equal(j, (a[1:] - a[:-1]) / 0.1)
    vd, ad, jd = planner.calculate_derivatives(torch.from_numpy(x).cuda(), 0.1)
    assert vd.is_cuda and np.allclose(jd.cpu().numpy(), j, rtol=0, atol=1e-9)
    planner.cleanup_gpu_memory()


def test_reference_planner_unit_cases():
    """The cases of the reference's tests/test_path_planning_unit.py:52-160, on a 2-joint planner:
    end points respected, batch = per-trajectory generator, positions clipped to the joint
    limits, torques clipped to the torque limits, shapes of the rollout and Cartesian outputs."""
    from manipulapy_b200 import ManipulatorDynamics, OptimizedTrajectoryPlanning
    from oracle import Oracle

    p = planar_2r_pack()
    dyn = ManipulatorDynamics(p["M"], S_list=p["S_list"], Glist=p["Glist"], Mlist_per_link=p["Mlist_per_link"])
    planner = OptimizedTrajectoryPlanning(dyn, "nonexistent.urdf", dyn, [(-1.0, 1.0), (-2.0, 2.0)])
    r = planner.joint_trajectory(thetastart=[0.0, 0.5], thetaend=[1.0, -0.5], Tf=1.0, N=4, method=3)
    assert r["positions"].shape == (4, 2)
    assert np.allclose(r["positions"][0], [0.0, 0.5]) and np.allclose(r["positions"][-1], [1.0, -0.5])
    sb = np.array([[0.0, 0.0], [0.5, -0.5]], dtype=np.float32)
    eb = np.array([[1.0, 1.0], [-0.5, 0.5]], dtype=np.float32)
    res = planner.batch_joint_trajectory(sb, eb, Tf=1.0, N=3, method=3)
    assert res["positions"].shape == (2, 3, 2)
    exp = Oracle.joint_trajectory(sb[0], eb[0], 1.0, 3, 3, np.array([[-1.0, 1.0], [-2.0, 2.0]]))
    assert np.array_equal(res["positions"][0], exp["positions"])
    planner1 = OptimizedTrajectoryPlanning(dyn, "nonexistent.urdf", dyn, [(-1.0, 1.0), (-1.0, 1.0)])
    pos = planner1.batch_joint_trajectory(np.zeros((1, 2), np.float32), np.array([[5.0, -5.0]], np.float32),
                                          Tf=1.0, N=4, method=3)["positions"]
    assert np.all(pos <= 1.0) and np.all(pos >= -1.0) and pos.max() == 1.0 and pos.min() == -1.0
    planner2 = OptimizedTrajectoryPlanning(dyn, "nonexistent.urdf", dyn, [(-1.0, 1.0)] * 2, [(-0.2, 0.2)] * 2)
    q = np.zeros((2, 2), np.float32)
    tq = planner2.inverse_dynamics_trajectory(q, np.zeros_like(q), np.ones_like(q) * 5.0)
    assert tq.dtype == np.float32 and np.all(tq <= np.float32(0.2)) and np.all(tq >= np.float32(-0.2))
    res = planner.forward_dynamics_trajectory(np.zeros(2, np.float32), np.zeros(2, np.float32), np.zeros((3, 2), np.float32),
                                              np.array([0, 0, -9.81], np.float32), np.zeros((3, 6), np.float32),
                                              dt=0.1, intRes=1)
    assert all(res[k].shape == (3, 2) and res[k].dtype == np.float32 for k in ("positions", "velocities", "accelerations"))
    Xs, Xe = np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)
    Xe[:3, 3] = [1.0, 0.0, 0.0]
    res = planner.cartesian_trajectory(Xs, Xe, Tf=1.0, N=5, method=3)
    assert res["positions"].shape == (5, 3) and res["orientations"].shape == (5, 3, 3)
    assert np.allclose(res["positions"][0], 0) and np.allclose(res["positions"][-1], [1, 0, 0])
    assert np.allclose(res["orientations"], np.eye(3))
    with pytest.raises(IndexError):  # zero steps (planning/trajectory_dynamics.py:612-615)
        planner.forward_dynamics_trajectory(np.zeros(2), np.zeros(2), np.zeros((0, 2)), [0, 0, -9.81], None, 0.1, 1)


def test_return_types_match_reference_api_contract(robots):
    """Return type / dtype / shape of every mirrored method on a 6-DOF robot, as pinned by the
    reference's tests/data/api_contract_golden.json (hot-path subset, tests/golden/api_contract.json)."""
    import json
    from pathlib import Path

    contract = json.loads((Path(__file__).resolve().parent / "golden" / "api_contract.json").read_text())
    rb = robots["ur5"]
    dyn, planner = rb.dynamics, rb.planner()
    rng = np.random.default_rng(0)
    th, dth, dd = rng.uniform(-1, 1, 6), rng.uniform(-1, 1, 6), rng.uniform(-1, 1, 6)
    g, ft = np.array([0, 0, -9.81]), np.zeros(6)
    N = 8
    traj = planner.joint_trajectory(th, th + 0.5, 2.0, N, 5)
    X0, X1 = dyn.forward_kinematics(th), dyn.forward_kinematics(th + 0.3)
    got = {
        "ManipulatorDynamics.forward_dynamics": dyn.forward_dynamics(th, dth, dd, g, ft),
        "ManipulatorDynamics.gravity_forces": dyn.gravity_forces(th),
        "ManipulatorDynamics.inverse_dynamics": dyn.inverse_dynamics(th, dth, dd, g, ft),
        "ManipulatorDynamics.mass_matrix": dyn.mass_matrix(th),
        "ManipulatorDynamics.velocity_quadratic_forces": dyn.velocity_quadratic_forces(th, dth),
        "OptimizedTrajectoryPlanning.cartesian_trajectory": planner.cartesian_trajectory(X0, X1, 2.0, N, 5),
        "OptimizedTrajectoryPlanning.forward_dynamics_trajectory":
            planner.forward_dynamics_trajectory(th, dth, rng.uniform(-5, 5, (N, 6)), g, np.zeros((N, 6)), 0.01, 1),
        "OptimizedTrajectoryPlanning.inverse_dynamics_trajectory":
            planner.inverse_dynamics_trajectory(traj["positions"], traj["velocities"], traj["accelerations"]),
        "OptimizedTrajectoryPlanning.joint_trajectory": traj,
        "SerialManipulator.forward_kinematics": dyn.forward_kinematics(th),
        "SerialManipulator.iterative_inverse_kinematics": dyn.iterative_inverse_kinematics(X1, th),
        "SerialManipulator.jacobian": dyn.jacobian(th),
        "SerialManipulator.smart_inverse_kinematics": dyn.smart_inverse_kinematics(X1, max_iterations=200),
        "SerialManipulator.robust_inverse_kinematics": dyn.robust_inverse_kinematics(X1, max_attempts=3, max_iterations=200),
        "SerialManipulator.end_effector_velocity": dyn.end_effector_velocity(th, dth),
    }
    # end_effector_velocity = J dtheta in either frame, batched = per row
    for frame in ("space", "body"):
        V = dyn.end_effector_velocity(np.stack([th, dd]), np.stack([dth, th]), frame)
        np.testing.assert_allclose(V[0], dyn.jacobian(th, frame) @ dth, rtol=0, atol=1e-14)
        np.testing.assert_allclose(V[1], dyn.jacobian(dd, frame) @ th, rtol=0, atol=1e-14)
    with pytest.raises(ValueError):
        dyn.end_effector_velocity(th, dth, "tool")

    def check(value, spec, where):
        if spec["type"] == "numpy.ndarray":
            assert isinstance(value, np.ndarray), where
            assert str(value.dtype) == spec["dtype"] and list(value.shape) == spec["shape"], (where, value.dtype, value.shape)
        elif spec["type"] == "dict":
            assert isinstance(value, dict) and set(value) == set(spec["items"]), where
            for k, sub in spec["items"].items():
                check(value[k], sub, f"{where}[{k}]")
        elif spec["type"] == "tuple":
            assert isinstance(value, tuple) and len(value) == len(spec["elements"]), where
            for i, sub in enumerate(spec["elements"]):
                check(value[i], sub, f"{where}[{i}]")
        else:
            assert type(value).__name__ == spec["type"], (where, type(value))

    for key, spec in contract.items():
        if not key.startswith("_"):
            check(got[key], spec["return"], key)


def test_registry_launcher_contract_vs_reference_golden():
    """`trajectory.*` registry launchers: the reference's registry contract (linear for other
    methods, N <= 1 / Tf <= 0 guards) against outputs of the reference's own launcher."""
    from manipulapy_b200 import KERNEL_REGISTRY

    g = load_golden("registry_trajectory")
    for name in ("linear", "method7", "cubic", "quintic", "n1", "tf0", "tfneg"):
        Tf, N, method = g[f"{name}_args"]
        got = KERNEL_REGISTRY.execute("trajectory.vectorized", g[f"{name}_start"], g[f"{name}_end"], float(Tf),
                                      int(N), int(method))
        for a, k in zip(got, ("positions", "velocities", "accelerations")):
            ref = g[f"{name}_{k}"]
            assert a.dtype == np.float32 and a.shape == ref.shape
            # (the reference's float32 polynomial carries a few float32 ulps of its largest term)
            np.testing.assert_allclose(a, ref, rtol=2e-6, atol=4e-6 * max(1.0, float(np.abs(ref).max(initial=0.0))),
                                       err_msg=f"{name} {k}")


def test_registry_launcher_runs_on_gpu():
    from manipulapy_b200 import execute_registered_kernel
    from oracle import Oracle

    s, e = np.array([0.1, -0.2, 0.3]), np.array([1.0, 0.5, -0.7])
    pos, vel, acc = execute_registered_kernel("trajectory.vectorized", s, e, 2.0, 64, 5)
    ref = Oracle.joint_trajectory(s, e, 2.0, 64, 5)
    assert _bits_equal(pos, ref["positions"]) and _bits_equal(vel, ref["velocities"]) and _bits_equal(acc, ref["accelerations"])


# ---------------------------------------------------------------------------------------------
# BASELINE.json full sizes: sampled oracle comparison + size-independent properties
# ---------------------------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------------
# float32-arithmetic kernel variants (north_star: 1e-4 relative on torques, 1e-5 on poses)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("robot", ROBOTS)
def test_float32_kernels_within_north_star_tolerance(robots, oracle_factory, robot):
    rb, o = robots[robot], oracle_factory(robot)
    n = rb.num_joints
    rng = np.random.default_rng(21)
    lo, hi = rb.joint_limits[:, 0], rb.joint_limits[:, 1]
    P = 4097  # ragged: not a multiple of the block size
    th = rng.uniform(lo, hi, (P, n))
    dth, ddth = rng.uniform(-2, 2, (P, n)), rng.uniform(-5, 5, (P, n))
    dyn = rb.dynamics
    ft = rng.uniform(-10, 10, 6)
    for Ftip in (None, ft):
        ref = o.inverse_dynamics(th[:600], dth[:600], ddth[:600], [0, 0, -9.81], Ftip, analytic=True)
        got = dyn.inverse_dynamics(th, dth, ddth, [0, 0, -9.81], Ftip, precision="float32")
        assert got.shape == (P, n) and got.dtype == np.float64
        assert _rel_rows(got[:600], ref) < 1e-4
        # and it really is the float32 kernel: it differs from the float64 one beyond float64 noise
        assert np.abs(got - dyn.inverse_dynamics(th, dth, ddth, [0, 0, -9.81], Ftip)).max() > 1e-9
    T, J = dyn.forward_kinematics_and_jacobian(th, precision="float32")
    assert T.dtype == np.float32 and J.dtype == np.float32 and T.shape == (P, 4, 4) and J.shape == (P, 6, n)
    Tref, Jref = o.forward_kinematics(th[:600]), o.jacobian(th[:600])
    assert np.abs(T[:600] - Tref).max() < 1e-5 * max(1.0, np.abs(Tref).max())
    assert np.abs(J[:600] - Jref).max() < 1e-5 * max(1.0, np.abs(Jref).max())
    np.testing.assert_array_equal(T[:, 3], np.tile(np.array([0, 0, 0, 1], np.float32), (P, 1)))
    # single-configuration calls keep their shapes
    assert dyn.forward_kinematics(th[0], precision="float32").shape == (4, 4)
    assert dyn.jacobian(th[0], precision=np.float32).shape == (6, n)
    with pytest.raises(ValueError):
        dyn.jacobian(th[0], precision="float16")


@pytest.mark.parametrize("robot", ["ur5", "iiwa14", "panda"])
def test_float32_fused_trajectory_inverse_dynamics(robots, oracle_factory, robot):
    """Trajectory rows stay bit-exact; float32 torques within 1e-4 of the float64 oracle; fused
    float32 kernel = float32 two-call sequence to float32 rounding; clipping still exact."""
    rb, o = robots[robot], oracle_factory(robot)
    n = rb.num_joints
    rng = np.random.default_rng(9)
    B, N = 23, 257
    s, e = rng.uniform(-3, 3, (B, n)), rng.uniform(-3, 3, (B, n))
    tl = np.array([[-60.0, 55.0]] * n)
    planner = rb.planner(torque_limits=tl)
    for ft in (None, [1.0, -2.0, 0.5, 3.0, 0.0, -1.0]):
        tau32, tr32 = planner.trajectory_inverse_dynamics(s, e, 2.0, N, 5, [0, 0, -9.81], ft, return_trajectory=True,
                                                          precision="float32")
        tau64, tr64 = planner.trajectory_inverse_dynamics(s, e, 2.0, N, 5, [0, 0, -9.81], ft, return_trajectory=True)
        assert all(_bits_equal(tr32[k], tr64[k]) for k in tr64)
        two32 = planner.inverse_dynamics_trajectory(tr64["positions"], tr64["velocities"], tr64["accelerations"],
                                                    [0, 0, -9.81], ft, precision="float32")
        # (float32 arithmetic: the fused kernel evaluates the joint rotations before the recursion and the
        # compiler contracts a few products differently than in the two-call kernel -- agreement to float32
        # rounding, not bit for bit; the float64 kernels ARE bit-identical, see
        # test_fused_trajectory_inverse_dynamics_equals_two_calls)
        assert _rel_rows(tau32.reshape(-1, n), two32.reshape(-1, n)) < 2e-6
        ref = o.inverse_dynamics_trajectory(tr64["positions"].reshape(-1, n), tr64["velocities"].reshape(-1, n),
                                            tr64["accelerations"].reshape(-1, n), [0, 0, -9.81], ft, tl, analytic=True)
        assert _rel_rows(tau32.reshape(-1, n), ref) < 1e-4
        assert tau32.max() <= np.float32(55.0) and tau32.min() >= np.float32(-60.0)
        assert not _bits_equal(tau32, tau64)


def test_cfg1_reference_case(robots, oracle_factory):
    """BASELINE config 1, the reference's own CPU-runnable case: UR5, seed 1, U(-1, 1)^6 end points,
    Tf = 2, N = 1000, quintic joint_trajectory followed by inverse_dynamics_trajectory.  Rows
    bit-exact against the reference golden; all 1000 torque rows against the oracle's LITERAL
    restatement of the reference algorithm (finite-difference Coriolis) at the reference's own
    golden tolerance, and against the analytic one to float32 rounding."""
    rb, o = robots["ur5"], oracle_factory("ur5")
    g = load_golden("trajectory")
    planner = rb.planner()
    tr = planner.joint_trajectory(g["cfg1_start"], g["cfg1_end"], 2.0, 1000, 5)
    for k in ("positions", "velocities", "accelerations"):
        assert _bits_equal(tr[k], g[f"cfg1_{k}"]), k
    tau = planner.inverse_dynamics_trajectory(tr["positions"], tr["velocities"], tr["accelerations"])
    assert tau.shape == (1000, 6) and tau.dtype == np.float32
    th, dth, dd = (tr[k].astype(np.float64) for k in ("positions", "velocities", "accelerations"))
    lit = o.inverse_dynamics_trajectory(th, dth, dd, analytic=False)
    np.testing.assert_allclose(tau, lit, rtol=1e-6, atol=1e-6)
    ana = o.inverse_dynamics_trajectory(th, dth, dd, analytic=True)
    np.testing.assert_allclose(tau, ana, rtol=3e-7, atol=1e-7)
    fused = planner.trajectory_inverse_dynamics(g["cfg1_start"], g["cfg1_end"], 2.0, 1000, 5)
    assert _bits_equal(fused, tau)


def test_full_size_cfg3_ur5_trajectory_rnea(robots, oracle_factory):
    """UR5, 4096 trajectories x 2441 steps (9,998,336 points), quintic, fused."""
    from oracle import Oracle

    rb, o = robots["ur5"], oracle_factory("ur5")
    rng = np.random.default_rng(3)
    B, N = 4096, 2441
    s = torch.from_numpy(rng.uniform(-np.pi, np.pi, (B, 6))).cuda()
    e = torch.from_numpy(rng.uniform(-np.pi, np.pi, (B, 6))).cuda()
    planner = rb.planner()
    tau, tr = planner.trajectory_inverse_dynamics(s, e, 2.0, N, 5, return_trajectory=True)
    assert tau.shape == (B, N, 6) and bool(torch.isfinite(tau).all())
    # (1) sampled trajectories against the oracle: rows bit-exact, torques to float32 rounding
    for b in (0, 1, 777, 4095):
        ref = Oracle.joint_trajectory(s[b].cpu().numpy()[None], e[b].cpu().numpy()[None], 2.0, N, 5, rb.joint_limits)
        for k in ref:
            assert _bits_equal(tr[k][b].cpu().numpy(), ref[k][0]), (b, k)
        rt = o.inverse_dynamics_trajectory(ref["positions"][0], ref["velocities"][0], ref["accelerations"][0], analytic=True)
        np.testing.assert_allclose(tau[b].cpu().numpy(), rt, rtol=3e-7, atol=1e-7)
    # (2) unfused two-kernel pipeline gives the same bits everywhere (checksum over all 6e7 values)
    two = planner.inverse_dynamics_trajectory(tr["positions"], tr["velocities"], tr["accelerations"])
    assert bool(torch.equal(two.view(torch.int32), tau.view(torch.int32)))
    # (3) endpoints: velocity / acceleration vanish, so torque = gravity torque at the end poses
    grav = rb.dynamics.gravity_forces(tr["positions"][:, 0].double())
    assert float((tau[:, 0].double() - grav).abs().max()) < 1e-4
    # (4) linearity of the torque in ddtheta at fixed (theta, dtheta): tau(a) - tau(0) = M a
    idx = torch.randint(0, B * N, (5000,), device="cuda")
    th = tr["positions"].reshape(-1, 6)[idx].double()
    dth = tr["velocities"].reshape(-1, 6)[idx].double()
    dd = tr["accelerations"].reshape(-1, 6)[idx].double()
    dyn = rb.dynamics
    lhs = dyn.inverse_dynamics(th, dth, dd, [0, 0, -9.81]) - dyn.inverse_dynamics(th, dth, torch.zeros_like(dd), [0, 0, -9.81])
    rhs = torch.einsum("pij,pj->pi", dyn.mass_matrix(th), dd)
    assert float((lhs - rhs).abs().max()) < 1e-9 * max(1.0, float(rhs.abs().max()))


def test_full_size_cfg5_billion_points(robots, oracle_factory):
    """BASELINE config 5 at its largest size on ONE GPU: 409,600 UR5 trajectories x 2441 steps =
    1.0e9 points (24 GB of float32 torques; element offsets beyond 2^32).  Size-independent checks:
    any slice of the big run is bit-identical to the same trajectories run as a small batch, the
    last trajectories (highest addresses) match the oracle, nothing is left unwritten."""
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs 40 GB of free device memory")
    rb, o = robots["ur5"], oracle_factory("ur5")
    planner = rb.planner()
    B, N = 409_600, 2441
    gen = torch.Generator(device="cuda").manual_seed(5)
    s = (torch.rand(B, 6, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1) * np.pi
    e = (torch.rand(B, 6, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1) * np.pi
    tau = planner.trajectory_inverse_dynamics(s, e, 2.0, N, 5)
    assert tau.shape == (B, N, 6) and tau.numel() > 2**32
    for lo in (0, 123_457, B - 1000):
        small = planner.trajectory_inverse_dynamics(s[lo:lo + 1000], e[lo:lo + 1000], 2.0, N, 5)
        assert bool(torch.equal(small.view(torch.int32), tau[lo:lo + 1000].view(torch.int32))), lo
        del small
    from oracle import Oracle
    for b in (B - 1, B // 2):
        ref = Oracle.joint_trajectory(s[b].cpu().numpy()[None], e[b].cpu().numpy()[None], 2.0, N, 5, rb.joint_limits)
        rt = o.inverse_dynamics_trajectory(ref["positions"][0], ref["velocities"][0], ref["accelerations"][0], analytic=True)
        np.testing.assert_allclose(tau[b].cpu().numpy(), rt, rtol=3e-7, atol=1e-7)
    # every chunk of the output was written with finite values
    for lo in range(0, B, 51_200):
        assert bool(torch.isfinite(tau[lo:lo + 51_200]).all())
    del tau
    torch.cuda.empty_cache()


def test_full_size_cfg2_million_fk_jacobian(robots, oracle_factory):
    """iiwa14 (true 7-DOF), the reference's 8-DOF Panda, and the 7-DOF Panda arm (its first seven
    screws; BASELINE config 2 says "Franka Panda 7-DOF"): 1,000,000 random configurations."""
    from manipulapy_b200 import load_robot
    from manipulapy_b200.robots import RobotBundle
    from oracle import Oracle

    p8 = load_robot("panda")
    panda7 = RobotBundle("panda7", p8.S_list[:, :7].copy(), p8.M, p8.Glist[:7], p8.Mlist_per_link[:7],
                         p8.joint_limits[:7])
    for robot in ("iiwa14", "panda", "panda7"):
        if robot == "panda7":
            rb, o = panda7, Oracle(panda7.S_list, panda7.M, panda7.Glist, panda7.Mlist_per_link)
        else:
            rb, o = robots[robot], oracle_factory(robot)
        n = rb.num_joints
        rng = np.random.default_rng(2)
        P = 1_000_000
        th = torch.from_numpy(rng.uniform(rb.joint_limits[:, 0], rb.joint_limits[:, 1], (P, n))).cuda()
        T, J = rb.dynamics.forward_kinematics_and_jacobian(th)
        assert T.shape == (P, 4, 4) and J.shape == (P, 6, n)
        R = T[:, :3, :3]
        eye = torch.eye(3, dtype=torch.float64, device="cuda")
        assert float((R @ R.transpose(1, 2) - eye).abs().max()) < 1e-13
        assert bool((T[:, 3] == torch.tensor([0.0, 0, 0, 1], dtype=torch.float64, device="cuda")).all())
        idx = rng.integers(0, P, 3000)
        thh = th[idx].cpu().numpy()
        assert np.abs(T[idx].cpu().numpy() - o.forward_kinematics(thh)).max() < 1e-12
        assert np.abs(J[idx].cpu().numpy() - o.jacobian(thh)).max() < 1e-12
        # finite-difference property: dp/dtheta_i = v_i + w_i x p for the space Jacobian
        eps = 1e-6
        k = 3
        thp, thm = th[:2000].clone(), th[:2000].clone()
        thp[:, k] += eps
        thm[:, k] -= eps
        dp = (rb.dynamics.forward_kinematics(thp)[:, :3, 3] - rb.dynamics.forward_kinematics(thm)[:, :3, 3]) / (2 * eps)
        Jk = J[:2000, :, k]
        pred = Jk[:, 3:] + torch.linalg.cross(Jk[:, :3], T[:2000, :3, 3])
        assert float((dp - pred).abs().max()) < 1e-8


@pytest.mark.parametrize("B", [65536, 8192], ids=["one_gpu", "eighth_share"])
def test_full_size_cfg4_iiwa_rollouts(robots, oracle_factory, B):
    """iiwa14, 65,536 rollouts x 1000 Euler steps (one warp per 32 rollouts), and one GPU's share of
    them on 8 GPUs (8,192: each step split across a warp pair); sampled rollouts against the oracle."""
    rb, o = robots["iiwa14"], oracle_factory("iiwa14")
    N, n = 1000, 7
    gen = torch.Generator(device="cuda").manual_seed(4)
    lo = torch.from_numpy(rb.joint_limits[:, 0]).cuda()
    hi = torch.from_numpy(rb.joint_limits[:, 1]).cuda()
    th0 = 0.5 * (lo + (hi - lo) * torch.rand(B, n, dtype=torch.float64, device="cuda", generator=gen))
    dth0 = torch.rand(B, n, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    # MPC-shooting style inputs: gravity-compensating nominal torque plus per-joint perturbations sized
    # to the link inertias.  (Pure U(-20, 20) N m noise on the iiwa wrist, inertia ~1e-3 kg m^2, drives
    # explicit Euler at dt = 1 ms to overflow for a few of 65,536 rollouts -- in the reference as well.)
    amp = torch.tensor([4.0, 4.0, 2.0, 2.0, 0.4, 0.2, 0.08], dtype=torch.float64, device="cuda")
    nominal = rb.dynamics.gravity_forces(th0)
    tau = (nominal[:, None, :] + (torch.rand(B, N, n, dtype=torch.float64, device="cuda", generator=gen) - 0.5) * amp).float()
    planner = rb.planner()
    r = planner.forward_dynamics_trajectory(th0, dth0, tau, [0, 0, -9.81], None, 1e-3, 1)
    pos, vel, acc = r["positions"], r["velocities"], r["accelerations"]
    assert pos.shape == (B, N, n)
    bad = torch.nonzero(~torch.isfinite(acc).all(dim=2).all(dim=1)).flatten().tolist()
    # explicit Euler may overflow on a rare rollout; then the oracle must overflow at the same step
    assert len(bad) <= 4, f"{len(bad)} rollouts with non-finite accelerations"
    for b in bad:
        rref = o.forward_dynamics_trajectory(th0[b].cpu().numpy(), dth0[b].cpu().numpy(), tau[b].double().cpu().numpy(),
                                             [0, 0, -9.81], None, 1e-3, 1, rb.joint_limits, analytic=True)
        fin_ref = np.isfinite(rref["accelerations"]).all(axis=1)
        fin_got = torch.isfinite(acc[b]).all(dim=1).cpu().numpy()
        first = int(np.argmin(fin_got))
        assert np.array_equal(fin_ref, fin_got), (
            f"rollout {b}: GPU non-finite from step {first}, oracle finite there: {bool(fin_ref[first])}; "
            f"state before: pos {pos[b, first - 1].tolist()} vel {vel[b, first - 1].tolist()} "
            f"acc {acc[b, first - 1].tolist()} tau {tau[b, first].tolist()}")
    assert bool((pos[:, 0] == th0.float()).all()) and bool((acc[:, 0] == 0).all())
    lo32, hi32 = lo.float(), hi.float()
    ok = torch.isfinite(pos)
    assert bool(((pos >= lo32) | ~ok).all()) and bool(((pos <= hi32) | ~ok).all())
    sel = [0, 1, 4097, B // 2 + 1, B - 7, B - 1]
    sel = [b for b in sel if b not in bad]
    # Against the oracle with the kernels' own LDL^T solve AND with the reference's LU: over all 1000 steps
    # the float32 rows agree to float32 rounding (measured: bit-identical for these rollouts,
    # scripts/cfg4_diag.py), so the solver is not a source of divergence at this step count.
    for solver in (2, True):
        ref = o.forward_dynamics_trajectory(th0[sel].cpu().numpy(), dth0[sel].cpu().numpy(),
                                            tau[sel].double().cpu().numpy(), [0, 0, -9.81], None, 1e-3, 1,
                                            rb.joint_limits, analytic=solver)
        for k, got in (("positions", pos), ("velocities", vel), ("accelerations", acc)):
            assert _rel_rows(got[sel].cpu().numpy().reshape(-1, n), ref[k].reshape(-1, n)) < 1e-6, (k, solver)


def test_full_size_cfg4_literal_torques(robots, oracle_factory):
    """SURVEY.md 8d cfg 4 to the letter: taumat ~ U(-20, 20) N m on every joint of the iiwa14, 65,536
    rollouts x 1000 steps.  20 N m on a wrist link of ~1e-3 kg m^2 is 2e4 rad/s^2: explicit Euler at
    dt = 1 ms leaves the joint range at once, the clip pins the angle, the velocity keeps integrating, and
    about 3 % of the rollouts overflow float32 before step 1000 -- in the reference's algorithm just the
    same.  Checked here: a rollout that overflows does so at the oracle's step; the rows before that, and
    all rows of the rollouts that stay finite, agree with the oracle to float32 rounding."""
    rb, o = robots["iiwa14"], oracle_factory("iiwa14")
    B, N, n = 65536, 1000, 7
    gen = torch.Generator(device="cuda").manual_seed(4)
    lo = torch.from_numpy(rb.joint_limits[:, 0]).cuda()
    hi = torch.from_numpy(rb.joint_limits[:, 1]).cuda()
    th0 = 0.5 * (lo + (hi - lo) * torch.rand(B, n, dtype=torch.float64, device="cuda", generator=gen))
    dth0 = torch.rand(B, n, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    tau = (torch.rand(B, N, n, dtype=torch.float32, device="cuda", generator=gen) - 0.5) * 40.0
    r = rb.planner().forward_dynamics_trajectory(th0, dth0, tau, [0, 0, -9.81], None, 1e-3, 1)
    pos, vel, acc = r["positions"], r["velocities"], r["accelerations"]
    fin = torch.isfinite(acc).all(dim=2) & torch.isfinite(vel).all(dim=2)
    bad = torch.nonzero(~fin.all(dim=1)).flatten().tolist()
    assert 0 < len(bad) < B // 10
    good = [b for b in (0, 1, 4097, B // 2 + 1, B - 1) if b not in bad][:4]
    sel = bad[:4] + good
    ref = o.forward_dynamics_trajectory(th0[sel].cpu().numpy(), dth0[sel].cpu().numpy(), tau[sel].double().cpu().numpy(),
                                        [0, 0, -9.81], None, 1e-3, 1, rb.joint_limits, analytic=2)
    with np.errstate(invalid="ignore", over="ignore"):
        for i, b in enumerate(sel):
            fin_ref = np.isfinite(ref["accelerations"][i]).all(1) & np.isfinite(ref["velocities"][i]).all(1)
            fin_got = fin[b].cpu().numpy()
            assert np.array_equal(fin_ref, fin_got), f"rollout {b}: overflow at different steps"
            upto = int(np.argmin(fin_got)) if not fin_got.all() else N
            assert (b in bad) == (upto < N)
            for k, got in (("positions", pos), ("velocities", vel), ("accelerations", acc)):
                assert _rel_rows(got[b, :upto].cpu().numpy(), ref[k][i, :upto]) < 1e-6, (b, k)
    lo32, hi32 = lo.float(), hi.float()
    okp = torch.isfinite(pos)
    assert bool(((pos >= lo32) | ~okp).all()) and bool(((pos <= hi32) | ~okp).all())


# ---------------------------------------------------------------------------------------------
# every robot of the reference's bundled database (oracle/gen_robot_zoo.py)
# ---------------------------------------------------------------------------------------------
from conftest import ZOO_ROBOTS, check_zoo_outputs, load_zoo  # noqa: E402


@pytest.mark.parametrize("robot", ZOO_ROBOTS)
def test_robot_zoo_vs_reference(robot):
    """26 URDF-derived chains (1 to 8 joints): FK, Jacobian, mass matrix, gravity / Coriolis forces,
    inverse and forward dynamics against outputs of the unmodified reference."""
    from manipulapy_b200 import ManipulatorDynamics

    z = load_zoo()[robot]
    dyn = ManipulatorDynamics(z["M"], None, None, None, z["S_list"], None, z["Glist"], z["Mlist_per_link"])
    th, dth, ddth, ft, g = z["thetas"], z["dthetas"], z["ddthetas"], z["ftips"], z["g"]
    check_zoo_outputs(
        z, dyn.forward_kinematics(th), dyn.jacobian(th), dyn.mass_matrix(th), dyn.gravity_forces(th, g),
        dyn.velocity_quadratic_forces(th, dth), dyn.inverse_dynamics(th, dth, ddth, g, ft),
        dyn.forward_dynamics(th, dth, z["taus"], g, ft))
    # a batch much larger than the golden rows gives the same bits for those rows
    rng = np.random.default_rng(1)
    n = th.shape[1]
    big = np.concatenate([th, rng.uniform(-1, 1, (997, n))])
    assert np.array_equal(dyn.mass_matrix(big)[: th.shape[0]], dyn.mass_matrix(th))


@pytest.mark.parametrize("robot,joints", [("ur10e", 2), ("abb_irb2400", 3), ("gen3", 4), ("xarm6", 5), ("crx10ia", 6),
                                          ("fanuc_lrmate", 6), ("abb_irb2400", 6), ("xarm6", 6), ("kinova_gen3", 7)])
def test_split_rollout_kernels_over_joint_counts(robot, joints):
    """The rollout kernels that split a step across warps, for every joint count they are compiled for (2 .. 7: chains
    cut from the robot database, general link geometry) and for the arm families with their own link-geometry
    kernels: 100 rollouts (three warps per 32 rollouts, ragged last block) against the oracle, and against the same
    rollouts inside batches that take the two-warp and the single-warp kernel (bit for bit)."""
    from manipulapy_b200 import ManipulatorDynamics, OptimizedTrajectoryPlanning
    from oracle import Oracle

    z = load_zoo()[robot]
    k = joints
    S, G, Mc, lim = z["S_list"][:, :k], z["Glist"][:k], z["Mlist_per_link"][:k], z["joint_limits"][:k]
    M = z["M"] if k == z["S_list"].shape[1] else Mc[-1]
    dyn = ManipulatorDynamics(M, None, None, None, S, None, G, Mc)
    planner = OptimizedTrajectoryPlanning(dyn, None, dyn, lim)
    o = Oracle(S, M, G, Mc)
    rng = np.random.default_rng(31 + k)
    B, N = 100, 30
    lo, hi = np.maximum(lim[:, 0], -3.0), np.minimum(lim[:, 1], 3.0)
    th0 = rng.uniform(0.5 * lo, 0.5 * hi, (B, k))
    dth0 = rng.uniform(-0.5, 0.5, (B, k))
    tau = rng.uniform(-3, 3, (B, N, k)).astype(np.float32)
    r = planner.forward_dynamics_trajectory(th0, dth0, tau, [0, 0, -9.81], None, 1e-3, 1)
    ref = o.forward_dynamics_trajectory(th0, dth0, tau.astype(np.float64), [0, 0, -9.81], None, 1e-3, 1, lim, analytic=True)
    for key in ref:
        assert _rel_rows(r[key].reshape(-1, k), ref[key].reshape(-1, k)) <= 1e-6, key
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for reps in (2 * 32 * sms // B + 1, 4 * 32 * sms // B + 1):  # two warps, one warp per 32 rollouts
        big = planner.forward_dynamics_trajectory(np.tile(th0, (reps, 1)), np.tile(dth0, (reps, 1)), np.tile(tau, (reps, 1, 1)),
                                                  [0, 0, -9.81], None, 1e-3, 1)
        for key in r:
            assert _bits_equal(big[key][:B], r[key]) and _bits_equal(big[key][-B:], r[key]), (key, reps)


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_singularity_callers(robot):
    """Singularity.condition_number / singularity_analysis / near_singularity_detection
    (singularity/singularity_analysis.py:52-74, 246-305) against the unmodified reference, one
    configuration at a time and batched; workspace_points = forward kinematics of the samples."""
    from manipulapy_b200 import SerialManipulator, Singularity

    g = load_golden("singularity")
    sm = SerialManipulator(M_list=g[f"{robot}_M"], S_list=g[f"{robot}_S"])
    sg = Singularity(sm)
    th, ref = g[f"{robot}_thetas"], g[f"{robot}_condition_number"]
    cond = sg.condition_number(th)
    reg = ref < 1e8  # regular configurations; the singular ones are rounding noise over ~1e-16
    np.testing.assert_allclose(cond[reg], ref[reg], rtol=1e-9)
    assert (cond[~reg] > 1e12).all() and (~reg).sum() == 2
    assert np.array_equal(sg.singularity_analysis(th), g[f"{robot}_singular"])
    assert np.array_equal(sg.near_singularity_detection(th), g[f"{robot}_near"])
    c3 = sg.condition_number(th[3])
    assert isinstance(c3, float) and abs(c3 - ref[3]) <= 1e-9 * ref[3]
    assert sg.singularity_analysis(th[0]) is True and sg.singularity_analysis(th[3]) is False
    lim = [(-1.0, 1.0)] * sm.num_joints
    pts, samples, hull = sg.workspace_points(lim, 20000, return_samples=True, return_hull=True)
    assert pts.shape == (20000, 3) and samples.dtype == np.float32 and np.abs(samples).max() <= 1.0
    np.testing.assert_allclose(pts, sm.forward_kinematics(samples.astype(np.float64))[:, :3, 3], rtol=0, atol=1e-12)
    assert hull.volume > 0 and np.array_equal(pts, sg.workspace_points(lim, 20000))  # seeded: reproducible


def test_widened_rows_edge_cases(robots):
    """Empty and single-row batches through the widened rows (IK modes and front ends, singularity
    callers, end-effector velocity), device-resident inputs, unknown options."""
    from manipulapy_b200 import Singularity

    from manipulapy_b200 import SerialManipulator

    rb = robots["ur5"]
    n = rb.num_joints
    sm = SerialManipulator(M_list=rb.M, S_list=rb.S_list, joint_limits=[tuple(r) for r in rb.joint_limits])
    none_T, none_th = np.empty((0, 4, 4)), np.empty((0, n))
    th, ok, it = sm.iterative_inverse_kinematics(none_T, none_th, adaptive_tuning=True, backtracking=True)
    assert th.shape == (0, n) and ok.shape == (0,) and it.shape == (0,)
    th, ok, it = sm.smart_inverse_kinematics(none_T)
    assert th.shape == (0, n) and ok.shape == (0,)
    th, ok, it, win = sm.robust_inverse_kinematics(none_T)
    assert th.shape == (0, n) and win.shape == (0,)
    assert sm.end_effector_velocity(none_th, none_th).shape == (0, 6)
    # one target as a batch of one keeps the batch axis
    T1 = sm.forward_kinematics(np.full((1, n), 0.3))
    th, ok, it = sm.smart_inverse_kinematics(T1, max_iterations=200)
    assert th.shape == (1, n) and ok.shape == (1,) and bool(ok[0])
    # device-resident targets
    thd, okd, itd = sm.iterative_inverse_kinematics(torch.from_numpy(T1).cuda(), torch.full((1, n), 0.25, dtype=torch.float64).cuda(),
                                                    backtracking=True)
    assert thd.is_cuda and okd.dtype == torch.bool and bool(okd[0])
    sg = Singularity(sm)
    c = sg.condition_number(torch.full((3, n), 0.3, dtype=torch.float64).cuda())
    assert c.is_cuda and c.shape == (3,) and bool((c > 1).all())
    assert sg.singularity_analysis(np.empty((0, n))).shape == (0,)
    assert sg.workspace_points([(-1, 1)] * n, 0).shape == (0, 3)
    with pytest.raises(NotImplementedError):
        sm.smart_inverse_kinematics(T1[0], strategy="cached", cache=object())


# ---------------------------------------------------------------------------------------------
# legacy dynamics path and its single-sample consumer, the computed-torque law
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_legacy_dynamics_path_vs_reference_golden(robot):
    """A ManipulatorDynamics built without Mlist_per_link, opted into the reference's legacy formulas:
    mass matrix and gravity to 1e-12, the finite-difference-based quantities to the finite-difference noise."""
    from manipulapy_b200 import ManipulatorDynamics

    g = load_golden("legacy_dynamics")
    S, M, G = g[f"{robot}_S"], g[f"{robot}_M"], g[f"{robot}_G"]
    with pytest.raises(NotImplementedError):
        ManipulatorDynamics(M, None, None, None, S, None, G)
    with pytest.warns(UserWarning):
        dyn = ManipulatorDynamics(M, None, None, None, S, None, G, legacy=True)
    th, dth, ddth, tau, ft, gv = (g[f"{robot}_{k}"] for k in ("th", "dth", "ddth", "tau", "ft", "g"))
    np.testing.assert_allclose(dyn.mass_matrix(th), g[f"{robot}_mass"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(dyn.mass_matrix(th[0]), g[f"{robot}_mass"][0], rtol=0, atol=1e-12)
    np.testing.assert_allclose(dyn.gravity_forces(th, gv), g[f"{robot}_grav"], rtol=0, atol=1e-12)
    # The Coriolis term is a central difference of mass matrices with eps = 1e-6: a last-bit difference in
    # one entry of M (the GPU's sin / cos, FMA contraction) is amplified by 1 / (2 eps) = 5e5 and summed over
    # n^2 entries times dtheta^2 -- the reference's numbers carry the same noise (a few 1e-7 here), so the
    # bar is per vector, relative to its largest entry.
    assert _rel_rows(dyn.velocity_quadratic_forces(th, dth), g[f"{robot}_cor"]) < 2e-6
    assert _rel_rows(dyn.inverse_dynamics(th, dth, ddth, gv, ft), g[f"{robot}_id"]) < 2e-6
    assert _rel_rows(dyn.forward_dynamics(th, dth, tau, gv, ft), g[f"{robot}_fd"]) < 2e-5
    # FK / Jacobian of such an object are the ordinary kinematics
    T = dyn.forward_kinematics(th)
    assert T.shape == (th.shape[0], 4, 4) and np.allclose(T[:, 3, 3], 1.0)
    # computed-torque law = M (Kp e + Ki eint + Kd de) + inverse dynamics at the desired acceleration
    thd, dthd, eint = th + 0.1, dth - 0.05, 0.01 * np.ones_like(th)
    Kp, Ki, Kd = 30.0, 0.5, 2.0
    got = dyn.computed_torque(thd, dthd, ddth, th, dth, gv, Kp, Ki, Kd, eint)
    want = (np.einsum("pij,pj->pi", g[f"{robot}_mass"], Kp * (thd - th) + Ki * eint + Kd * (dthd - dth))
            + dyn.inverse_dynamics(th, dth, ddth, gv, None))
    assert _rel_rows(got, want) < 2e-6


def test_computed_torque_on_the_per_link_model(robots, oracle_factory):
    rb, o = robots["ur5"], oracle_factory("ur5")
    rng = np.random.default_rng(17)
    th, dth, ddthd = rng.uniform(-2, 2, (50, 6)), rng.uniform(-1, 1, (50, 6)), rng.uniform(-2, 2, (50, 6))
    thd, dthd = th + rng.uniform(-0.1, 0.1, (50, 6)), dth + rng.uniform(-0.1, 0.1, (50, 6))
    Kp, Kd = np.full(6, 40.0), np.linspace(1, 3, 6)
    got = rb.dynamics.computed_torque(thd, dthd, ddthd, th, dth, [0, 0, -9.81], Kp, 0.0, Kd)
    M = o.mass_matrix(th, analytic=True)
    want = (np.einsum("pij,pj->pi", M, Kp * (thd - th) + Kd * (dthd - dth))
            + o.inverse_dynamics(th, dth, ddthd, analytic=True))
    assert _rel_rows(got, want) < 1e-9
