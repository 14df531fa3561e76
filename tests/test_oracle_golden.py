"""Pin the CPU oracle (oracle/oracle.c) against the reference.

* the reference's own golden vectors (tests/data/dynamics_golden_{ur5,panda}.npz, replayed
  verbatim inside tests/golden/dynamics_*.npz) at the reference's own tolerances
  (tests/test_dynamics_golden.py:77-83: rtol 1e-7, atol 1e-9 for M/g, 1e-8 for c/ID);
* outputs of the unmodified reference run by oracle/gen_golden.py (FK, Jacobian, forward
  dynamics, joint/batch trajectories, inverse/forward dynamics trajectories);
* the analytic known answers the reference tests use (2R planar arm).
"""

import numpy as np
import pytest

from conftest import load_golden, load_pack, planar_2r_pack

RTOL = 1e-7
ATOL = {"mass_matrix": 1e-9, "gravity_forces": 1e-9,
        "velocity_quadratic_forces": 1e-8, "inverse_dynamics": 1e-8}
ROBOTS = ["ur5", "panda", "iiwa14"]


@pytest.mark.parametrize("analytic", [False, True], ids=["literal", "analytic"])
@pytest.mark.parametrize("robot", ROBOTS)
def test_dynamics_golden(oracle_factory, robot, analytic):
    o = oracle_factory(robot)
    g = load_golden(f"dynamics_{robot}")
    th, dth, ddth, ft = g["thetas"], g["dthetas"], g["ddthetas"], g["ftips"]
    got = {
        "mass_matrix": o.mass_matrix(th, analytic),
        "gravity_forces": o.gravity_forces(th, g["g"], analytic),
        "velocity_quadratic_forces": o.velocity_quadratic_forces(th, dth, analytic),
        "inverse_dynamics": o.inverse_dynamics(th, dth, ddth, g["g"], ft, analytic),
    }
    for k, v in got.items():
        np.testing.assert_allclose(v, g[k], rtol=RTOL, atol=ATOL[k], err_msg=k)


@pytest.mark.parametrize("robot", ROBOTS)
def test_kinematics_vs_reference(oracle_factory, robot):
    o = oracle_factory(robot)
    g = load_golden(f"dynamics_{robot}")
    np.testing.assert_allclose(o.forward_kinematics(g["thetas"]), g["forward_kinematics"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(o.jacobian(g["thetas"]), g["jacobian"], rtol=0, atol=1e-14)


@pytest.mark.parametrize("analytic", [False, True], ids=["literal", "analytic"])
@pytest.mark.parametrize("robot", ROBOTS)
def test_forward_dynamics_vs_reference(oracle_factory, robot, analytic):
    o = oracle_factory(robot)
    g = load_golden(f"dynamics_{robot}")
    i = g["fd_index"]
    got = o.forward_dynamics(g["thetas"][i], g["dthetas"][i], g["fd_tau"], g["g"], g["ftips"][i], analytic)
    ref = g["forward_dynamics"]
    # accelerations reach 1e5 rad/s^2 (lambda_min(M) ~ 1e-4): per-vector inf-norm relative
    scale = np.maximum(1.0, np.abs(ref).max(axis=1, keepdims=True))
    assert np.max(np.abs(got - ref) / scale) < 1e-9


def _assert_traj_equal(got, ref, tag):
    """Bit-exact float32, except at exact-cancellation points (tau = 0.5 or 1) where the
    reference's Numba ``fastmath=True`` kernel (planning/trajectory.py:14) contracts the
    polynomial into FMAs and leaves O(1e-16) residues instead of the IEEE zero."""
    assert got.dtype == np.float32 and got.shape == ref.shape, tag
    neq = got.view(np.uint32) != ref.view(np.uint32)
    assert neq.mean() <= 0.5 and np.all(np.abs(got[neq].astype(np.float64) - ref[neq]) <= 1e-12), tag
    assert np.all(got[neq] == 0), tag


def test_joint_trajectory_bit_exact_vs_reference():
    from oracle import Oracle

    g = load_golden("trajectory")
    lim = g["joint_limits"]
    for name in ("cfg1", "cubic50", "two", "clipped", "odd_tf"):
        Tf, N, method = g[f"{name}_args"]
        r = Oracle.joint_trajectory(g[f"{name}_start"], g[f"{name}_end"], Tf, int(N), int(method), lim)
        for k in ("positions", "velocities", "accelerations"):
            _assert_traj_equal(r[k], g[f"{name}_{k}"], (name, k))
        if name in ("cfg1", "odd_tf"):  # no exact-cancellation rows: every bit equal
            assert all(np.array_equal(r[k].view(np.uint32), g[f"{name}_{k}"].view(np.uint32)) for k in r)
    Tf, N, method = g["batch_args"]
    r = Oracle.joint_trajectory(g["batch_start"], g["batch_end"], Tf, int(N), int(method), lim)
    for k in ("positions", "velocities", "accelerations"):
        _assert_traj_equal(r[k], g[f"batch_{k}"], ("batch", k))


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_inverse_dynamics_trajectory_vs_reference(oracle_factory, robot):
    o = oracle_factory(robot)
    g = load_golden("id_trajectory")
    th, dth, ddth = g[f"{robot}_theta"], g[f"{robot}_dtheta"], g[f"{robot}_ddtheta"]
    tl = g[f"{robot}_torque_limits"]
    for analytic in (False, True):
        t0 = o.inverse_dynamics_trajectory(th, dth, ddth, torque_limits=tl, analytic=analytic)
        t1 = o.inverse_dynamics_trajectory(th, dth, ddth, g[f"{robot}_g1"], g[f"{robot}_ftip1"], tl, analytic)
        for got, ref in ((t0, g[f"{robot}_tau_default"]), (t1, g[f"{robot}_tau_g1_ftip1"])):
            assert got.dtype == np.float32
            # float32 outputs: equal up to the last float32 ulp (finite-difference noise ~1e-9
            # can flip a rounding); clipped entries must be exactly the limit
            np.testing.assert_allclose(got, ref, rtol=3e-7, atol=1e-7)
            assert np.array_equal(got == tl[:, 0].astype(np.float32), ref == tl[:, 0].astype(np.float32))
            assert np.array_equal(got == tl[:, 1].astype(np.float32), ref == tl[:, 1].astype(np.float32))
        if robot == "ur5":  # the finite torque limits really clip on this fixture
            assert (t0 == np.float32(35.0)).any() or (t0 == np.float32(-40.0)).any()


def _assert_rows_close(got, ref, tol, tag):
    """Per-row inf-norm relative: |d| <= tol * max(1, ||ref_row||_inf).  Accelerations reach
    3e5 rad/s^2 on these fixtures (lambda_min(M) ~ 1e-4), so element-wise relative is meaningless."""
    assert got.dtype == ref.dtype and got.shape == ref.shape, tag
    scale = np.maximum(1.0, np.abs(ref).max(axis=-1, keepdims=True))
    assert np.max(np.abs(got.astype(np.float64) - ref) / scale) <= tol, tag


@pytest.mark.parametrize("analytic", [False, True], ids=["literal", "analytic"])
@pytest.mark.parametrize("robot", ["iiwa14", "ur5"])
def test_forward_dynamics_trajectory_vs_reference(oracle_factory, robot, analytic):
    o = oracle_factory(robot)
    g = load_golden("fd_trajectory")
    lim = g[f"{robot}_joint_limits"]
    dt, intres = g[f"{robot}_a_args"]
    r = o.forward_dynamics_trajectory(g[f"{robot}_a_theta0"], g[f"{robot}_a_dtheta0"], g[f"{robot}_a_tau"],
                                      [0, 0, -9.81], None, dt, int(intres), lim, analytic)
    for k in ("positions", "velocities", "accelerations"):
        _assert_rows_close(r[k], g[f"{robot}_a_{k}"], 1e-6, k)
    dt, intres = g[f"{robot}_b_args"]
    r = o.forward_dynamics_trajectory(g[f"{robot}_b_theta0"], g[f"{robot}_b_dtheta0"], g[f"{robot}_b_tau"],
                                      g[f"{robot}_b_g"], g[f"{robot}_b_ftip"], dt, int(intres), lim, analytic)
    ref_pos = g[f"{robot}_b_positions"]
    for k in ("positions", "velocities", "accelerations"):
        _assert_rows_close(r[k], g[f"{robot}_b_{k}"], 1e-6, k)
    # the joint-limit clip was exercised and lands exactly on the float32 limit
    hi32 = lim[:, 1].astype(np.float32)
    assert (ref_pos[1:] == hi32).any()
    assert np.array_equal(r["positions"] == hi32, ref_pos == hi32)


def test_forward_dynamics_trajectory_zero_steps(oracle_factory):
    o = oracle_factory("ur5")
    with pytest.raises(IndexError):
        o.forward_dynamics_trajectory(np.zeros(6), np.zeros(6), np.zeros((0, 6)), [0, 0, -9.81], None, 1e-3, 1)


def test_planar_2r_known_answers():
    """Murray-Li-Sastry Ex. 4.3 mass matrix and hand-derived holding torques
    (reference tests/test_v132_regressions.py:126-192, 229-286)."""
    from oracle import Oracle

    p = planar_2r_pack()
    o = Oracle(p["S_list"], p["M"], p["Glist"], p["Mlist_per_link"])
    th = np.array([0.0, np.pi / 2])
    c2 = np.cos(th[1])
    Mexp = np.array([[1 + (1 + 1 + 2 * c2), 1 + c2], [1 + c2, 1.0]])
    for analytic in (False, True):
        np.testing.assert_allclose(o.mass_matrix(th, analytic)[0], Mexp, atol=1e-12)
        np.testing.assert_allclose(o.gravity_forces(th, [-9.81, 0, 0], analytic)[0], [-9.81, -9.81], atol=1e-12)
        np.testing.assert_allclose(o.gravity_forces(th, [0, -9.81, 0], analytic)[0], [19.62, 0.0], atol=1e-12)
        np.testing.assert_allclose(o.gravity_forces(th, [0, 0, -9.81], analytic)[0], [0.0, 0.0], atol=1e-12)


def test_quintic_endpoints_and_linear_contract():
    """Endpoint accelerations vanish for quintic (reference tests/test_v132_regressions.py:516-541)."""
    from oracle import Oracle

    r = Oracle.joint_trajectory(np.zeros(3), np.ones(3), 2.0, 11, 5)
    assert np.all(r["accelerations"][0] == 0) and np.allclose(r["accelerations"][-1], 0, atol=1e-6)
    assert np.all(r["positions"][0] == 0) and np.allclose(r["positions"][-1], 1.0)
    assert np.all(r["velocities"][0] == 0)


def test_registry_trajectory_restatement_bit_exact_vs_reference():
    """Oracle.registry_trajectory against the reference's own registry launcher with CUDA
    routing off (tests/golden/registry_trajectory.npz, oracle/gen_golden.py): linear scaling for
    other methods, N <= 1 / Tf <= 0 guards.  Same float32 NumPy arithmetic: bit-exact."""
    from oracle import Oracle

    g = load_golden("registry_trajectory")
    for name in ("linear", "method7", "cubic", "quintic", "n1", "tf0", "tfneg"):
        Tf, N, method = g[f"{name}_args"]
        got = Oracle.registry_trajectory(g[f"{name}_start"], g[f"{name}_end"], float(Tf), int(N), int(method))
        for a, k in zip(got, ("positions", "velocities", "accelerations")):
            ref = g[f"{name}_{k}"]
            assert a.dtype == np.float32 and a.shape == ref.shape
            assert np.array_equal(a.view(np.uint32), ref.view(np.uint32)), (name, k)


def test_body_frame_restatement_vs_reference():
    """Oracle.body_forward_kinematics / body_jacobian against the unmodified reference
    (frame="body"; tests/golden/body_kinematics.npz), including a chain whose B_list is
    independent of its S_list."""
    from oracle import Oracle

    g = load_golden("body_kinematics")
    for k in ("ur5", "free"):
        T = Oracle.body_forward_kinematics(g[f"{k}_M"], g[f"{k}_B"], g[f"{k}_theta"])
        J = Oracle.body_jacobian(g[f"{k}_B"], g[f"{k}_theta"])
        np.testing.assert_allclose(T, g[f"{k}_T"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(J, g[f"{k}_J"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_inverse_kinematics_restatement_vs_reference(robot):
    """Oracle.iterative_inverse_kinematics against the unmodified reference
    (tests/golden/inverse_kinematics.npz): same iterates, success flags and iteration counts --
    including the runs that go through the stagnation restart, because both draw it from
    NumPy's global generator seeded alike."""
    from oracle import Oracle

    g = load_golden("inverse_kinematics")
    pack = load_pack(robot)
    o = Oracle(g[f"{robot}_S"], g[f"{robot}_M"], pack["Glist"], pack["Mlist_per_link"])  # the golden's own M
    lim = [tuple(r) for r in g[f"{robot}_limits"]]
    for i, (Td, seed, par) in enumerate(zip(g[f"{robot}_T"], g[f"{robot}_seed"], g[f"{robot}_params"])):
        np.random.seed(100 + i)
        th, ok, it = o.iterative_inverse_kinematics(Td, seed, max_iterations=int(par[0]), damping=par[1], step_cap=par[2],
                                                    weight_orientation=par[3], weight_position=par[4], joint_limits=lim)
        assert ok == bool(g[f"{robot}_success"][i]), i
        assert it == int(g[f"{robot}_iterations"][i]), i
        np.testing.assert_allclose(th, g[f"{robot}_theta"][i], rtol=0, atol=1e-7, err_msg=str(i))


IK_MODES = {"adaptive": dict(adaptive_tuning=True), "backtracking": dict(backtracking=True),
            "both": dict(adaptive_tuning=True, backtracking=True)}


@pytest.mark.parametrize("mode", sorted(IK_MODES))
@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_inverse_kinematics_modes_restatement_vs_reference(robot, mode):
    """... and with the Levenberg-Marquardt adaptation and / or the line search switched on
    (tests/golden/inverse_kinematics_modes.npz, generated from the unmodified reference)."""
    from oracle import Oracle

    g = load_golden("inverse_kinematics_modes")
    pack = load_pack(robot)
    o = Oracle(g[f"{robot}_S"], g[f"{robot}_M"], pack["Glist"], pack["Mlist_per_link"])
    lim = [tuple(r) for r in g[f"{robot}_limits"]]
    for i, (Td, seed, budget) in enumerate(zip(g[f"{robot}_T"], g[f"{robot}_seed"], g[f"{robot}_max_iterations"])):
        np.random.seed(200 + i)
        th, ok, it = o.iterative_inverse_kinematics(Td, seed, max_iterations=int(budget), joint_limits=lim,
                                                    **IK_MODES[mode])
        assert ok == bool(g[f"{robot}_{mode}_success"][i]), i
        assert it == int(g[f"{robot}_{mode}_iterations"][i]), i
        np.testing.assert_allclose(th, g[f"{robot}_{mode}_theta"][i], rtol=0, atol=1e-7, err_msg=str(i))


def _oracle_ik_solver(o, lim, **fixed):
    def solve(Tds, th0, **kw):
        r = [o.iterative_inverse_kinematics(T, t, joint_limits=lim, adaptive_tuning=True, backtracking=True,
                                            **fixed, **kw) for T, t in zip(Tds, th0)]
        return np.stack([x[0] for x in r]), np.array([x[1] for x in r]), np.array([x[2] for x in r])
    return solve


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_ik_front_end_drivers_vs_reference(robot):
    """The host-side restart logic of smart_ / robust_inverse_kinematics (manipulapy_b200.ik_helpers:
    initial guesses, fall-back rounds, best-iterate tracking, NumPy-generator draws) with the oracle's
    solver plugged in reproduces the unmodified reference on every golden case -- success, total
    iterations, winning strategy, solution -- including those with fall-back starts and restarts."""
    from manipulapy_b200 import ik_helpers
    from oracle import Oracle

    g = load_golden("inverse_kinematics_front_ends")
    pack = load_pack(robot)
    o = Oracle(g[f"{robot}_S"], g[f"{robot}_M"], pack["Glist"], pack["Mlist_per_link"])
    lim = [tuple(r) for r in g[f"{robot}_limits"]]
    n = len(lim)
    fk = o.forward_kinematics
    for i, Td in enumerate(g[f"{robot}_T"]):
        np.random.seed(300 + i)
        th, ok, it = ik_helpers.smart_driver(_oracle_ik_solver(o, lim, max_iterations=120), fk, Td[None], n, lim,
                                             "workspace_heuristic", True)
        assert bool(ok[0]) == bool(g[f"{robot}_smart_success"][i]) and int(it[0]) == int(g[f"{robot}_smart_iterations"][i]), i
        np.testing.assert_allclose(th[0], g[f"{robot}_smart_theta"][i], rtol=0, atol=1e-9, err_msg=str(i))
        np.random.seed(400 + i)
        th, ok, it, win = ik_helpers.robust_driver(_oracle_ik_solver(o, lim, eomg=2e-3, ev=2e-3, max_iterations=120),
                                                   fk, Td[None], n, lim, 4)
        assert bool(ok[0]) == bool(g[f"{robot}_robust_success"][i]) and int(it[0]) == int(g[f"{robot}_robust_iterations"][i]), i
        assert str(win[0]) == str(g[f"{robot}_robust_strategy"][i]), i
        np.testing.assert_allclose(th[0], g[f"{robot}_robust_theta"][i], rtol=0, atol=1e-9, err_msg=str(i))


CART_CASES = ("generic5", "generic3", "method1", "same_R", "tiny", "near_pi", "pi_band", "pi_exact")


def test_cartesian_trajectory_restatement_vs_reference():
    """Oracle.cartesian_trajectory against the unmodified reference
    (tests/golden/cartesian_trajectory.npz): float32 outputs equal to one float32 rounding,
    including rotations inside MatrixLog3's half-turn band."""
    from oracle import Oracle

    g = load_golden("cartesian_trajectory")
    for name in CART_CASES:
        Tf, N, method = g[f"{name}_args"]
        got = Oracle.cartesian_trajectory(g[f"{name}_Xstart"], g[f"{name}_Xend"], float(Tf), int(N), int(method))
        for k in ("positions", "velocities", "accelerations", "orientations"):
            assert got[k].dtype == np.float32 and got[k].shape == g[f"{name}_{k}"].shape
            np.testing.assert_allclose(got[k], g[f"{name}_{k}"], rtol=3e-7, atol=1e-7, err_msg=f"{name} {k}")


# ---------------------------------------------------------------------------------------------
# every robot of the reference's bundled database (oracle/gen_robot_zoo.py)
# ---------------------------------------------------------------------------------------------
from conftest import ZOO_ROBOTS, check_zoo_outputs, load_zoo  # noqa: E402


@pytest.mark.parametrize("analytic", [False, True], ids=["literal", "analytic"])
@pytest.mark.parametrize("robot", ZOO_ROBOTS)
def test_robot_zoo_vs_reference(robot, analytic):
    from oracle import Oracle

    z = load_zoo()[robot]
    o = Oracle(z["S_list"], z["M"], z["Glist"], z["Mlist_per_link"])
    th, dth, ddth, ft, g = z["thetas"], z["dthetas"], z["ddthetas"], z["ftips"], z["g"]
    check_zoo_outputs(
        z, o.forward_kinematics(th), o.jacobian(th), o.mass_matrix(th, analytic),
        o.gravity_forces(th, g, analytic), o.velocity_quadratic_forces(th, dth, analytic),
        o.inverse_dynamics(th, dth, ddth, g, ft, analytic),
        o.forward_dynamics(th, dth, z["taus"], g, ft, analytic))


# ---------------------------------------------------------------------------------------------
# collision / limit post-processing hook of joint_trajectory
# ---------------------------------------------------------------------------------------------
def _collision_case(robot):
    from oracle.oracle_lib import CollisionOracle

    g = load_golden("collision")
    pack = load_pack(robot)
    from conftest import ROBOTS as ROBOT_DIR

    with np.load(ROBOT_DIR / f"{robot}_links.npz") as d:
        links = {k: d[k] for k in d.files}
    names = [str(x) for x in links["link_names"]]
    hulls = {names.index(str(nm)): g[f"{robot}_hull_{nm}"] for nm in g[f"{robot}_hull_links"]}
    return g, links, CollisionOracle(pack["S_list"], links["link_joint"], links["link_home"], hulls, links["link_acm"])


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_collision_oracle_vs_reference_golden(robot):
    """Every link pose of URDF.link_fk_batch, the checker's flags and the rows the hook nudged, as
    recorded from the unmodified reference with injected hulls."""
    g, links, co = _collision_case(robot)
    cfgs = g[f"{robot}_cfgs"]
    np.testing.assert_allclose(co.link_fk_batch(cfgs), g[f"{robot}_link_fk"], rtol=0, atol=1e-12)
    assert np.array_equal(co.check_collision(cfgs), g[f"{robot}_flags"])
    assert 0 < g[f"{robot}_flags"].sum() < len(cfgs)
    for k in (0, 1):
        t = f"{robot}_traj{k}_"
        raw, end = g[t + "raw"], g[t + "end"].astype(np.float32)
        assert np.array_equal(co.check_collision(raw), g[t + "flags_before"])
        got, iters = co.avoid(raw, end, attractive_gain=g[f"{robot}_gains"][0])
        assert np.array_equal(got.view(np.uint32), g[t + "positions"].view(np.uint32))  # float32 steps: bit for bit
        assert np.array_equal(co.check_collision(got), g[t + "flags_after"])
        assert (iters[g[t + "flags_before"] == 0] == 0).all()
        if k == 1:  # ends in collision: rows run into the 100-iteration cap
            assert (iters == 100).any() and g[t + "flags_after"].any()


# ---------------------------------------------------------------------------------------------
# legacy dynamics path: objects built without Mlist_per_link
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_legacy_oracle_vs_reference_golden(robot):
    from oracle.oracle_lib import LegacyOracle

    g = load_golden("legacy_dynamics")
    lo = LegacyOracle(g[f"{robot}_S"], g[f"{robot}_M"], g[f"{robot}_G"])
    for p in range(g[f"{robot}_th"].shape[0]):
        th, dth, ddth = g[f"{robot}_th"][p], g[f"{robot}_dth"][p], g[f"{robot}_ddth"][p]
        np.testing.assert_allclose(lo.mass_matrix(th), g[f"{robot}_mass"][p], rtol=0, atol=1e-12)
        np.testing.assert_allclose(lo.gravity_forces(th, g[f"{robot}_g"]), g[f"{robot}_grav"][p], rtol=0, atol=1e-12)
        # finite differences of eps = 1e-6: the reference's own noise floor is ~1e-9
        np.testing.assert_allclose(lo.velocity_quadratic_forces(th, dth), g[f"{robot}_cor"][p], rtol=1e-7, atol=1e-8)
        np.testing.assert_allclose(lo.inverse_dynamics(th, dth, ddth, g[f"{robot}_g"], g[f"{robot}_ft"][p]),
                                   g[f"{robot}_id"][p], rtol=1e-7, atol=1e-8)
        fd = lo.forward_dynamics(th, dth, g[f"{robot}_tau"][p], g[f"{robot}_g"], g[f"{robot}_ft"][p])
        assert np.abs(fd - g[f"{robot}_fd"][p]).max() <= 1e-6 * max(1.0, np.abs(g[f"{robot}_fd"][p]).max())
