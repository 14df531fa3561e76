"""CPU-side checks of the native path (no GPU needed).

* the C-ABI library loads and exports every symbol ``include/mpk.h`` declares;
* host-only entry points (robot pack construction) behave and report errors;
* the kernels' own per-thread templates, executed on the CPU through tests/hostcheck,
  reproduce the reference's golden vectors and the oracle -- the algebra the GPU will run
  is pinned before any GPU time is spent;
* host logic: operator registry contract, sharding under gloo (world_size 2), and that the
  product path refuses to run without CUDA instead of falling back.
"""

import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from pathlib import Path as _Path

GOLDEN_DIR = _Path(__file__).resolve().parent / "golden"

from conftest import REPO, load_golden, load_pack, planar_2r_pack, random_general_pack

ROBOTS = ["ur5", "panda", "iiwa14", "xarm6"]


def test_abi_exports_every_declared_symbol():
    from manipulapy_b200 import _native

    lib = _native.lib()
    syms = _native.declared_symbols()
    assert len(syms) >= 14 and "mpk_inverse_dynamics" in syms and "mpk_forward_dynamics_trajectory" in syms
    for s in syms:
        assert hasattr(lib, s), f"libmpk.so does not export {s}"
    assert lib.mpk_version() >= 100


def test_torch_ops_register():
    from manipulapy_b200 import _native

    ops = _native.ops()
    for name in ("robot_create", "joint_trajectory", "fk_jacobian", "inverse_dynamics",
                 "trajectory_inverse_dynamics", "mass_matrix", "forward_dynamics",
                 "forward_dynamics_trajectory", "fma_peak"):
        assert hasattr(ops, name)


def test_robot_create_errors(hostcheck):
    p = load_pack("ur5")
    bad = dict(p)
    bad["S_list"] = p["S_list"].copy()
    bad["S_list"][:3, 2] *= 1.5  # non-unit omega: the reference formula is not a rigid motion
    with pytest.raises(RuntimeError, match="non-unit"):
        hostcheck.robot(bad)
    hel = dict(p)
    hel["S_list"] = p["S_list"].copy()
    hel["S_list"][3:, 1] += 0.1 * hel["S_list"][:3, 1]  # pitch: omega . v != 0
    with pytest.raises(RuntimeError, match="helical"):
        hostcheck.robot(hel)
    big = random_general_pack(8, 0)
    big9 = {k: np.concatenate([v, v[..., :1]], -1) if k == "S_list" else v for k, v in big.items()}
    big9["Glist"] = np.concatenate([big["Glist"], big["Glist"][:1]])
    big9["Mlist_per_link"] = np.concatenate([big["Mlist_per_link"], big["Mlist_per_link"][:1]])
    with pytest.raises(RuntimeError, match="1..8"):
        hostcheck.robot(big9)
    h, n = hostcheck.robot(p)
    assert hostcheck.L.mpk_robot_dof(h) == 6 and hostcheck.L.mpk_robot_is_rigid(h) == 1
    h2, _ = hostcheck.robot(p, flags=1)
    assert hostcheck.L.mpk_robot_is_rigid(h2) == 0


@pytest.mark.parametrize("flags", [0, 1], ids=["rigid", "general"])
@pytest.mark.parametrize("robot", ["ur5", "panda", "iiwa14"])
def test_kernel_algebra_vs_reference_golden(hostcheck, robot, flags):
    """Reference tolerances (tests/test_dynamics_golden.py:77-83): rtol 1e-7, atol 1e-9 / 1e-8."""
    g = load_golden(f"dynamics_{robot}")
    rb = hostcheck.robot(load_pack(robot), flags)
    th, dth, ddth = g["thetas"], g["dthetas"], g["ddthetas"]
    T, J = hostcheck.fk(rb, th)
    np.testing.assert_allclose(T, g["forward_kinematics"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(J, g["jacobian"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(hostcheck.mass(rb, th), g["mass_matrix"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(hostcheck.rnea(rb, th, g=g["g"]), g["gravity_forces"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(hostcheck.rnea(rb, th, dth, g=(0, 0, 0)), g["velocity_quadratic_forces"],
                               rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(hostcheck.rnea(rb, th, dth, ddth, g["g"], g["ftips"]), g["inverse_dynamics"],
                               rtol=1e-7, atol=1e-8)
    i = g["fd_index"]
    dd = hostcheck.fd(rb, th[i], dth[i], g["fd_tau"], g["g"], g["ftips"][i])
    ref = g["forward_dynamics"]
    assert np.max(np.abs(dd - ref) / np.maximum(1.0, np.abs(ref).max(1, keepdims=True))) < 1e-9


@pytest.mark.parametrize("robot", ROBOTS)
def test_kernel_algebra_vs_oracle_random(hostcheck, oracle_factory, robot):
    """Against the analytic oracle the only difference is rounding: 1e-9 * max(1, |ref|_inf)."""
    p = load_pack(robot)
    o = oracle_factory(robot)
    n = p["S_list"].shape[1]
    rng = np.random.default_rng(11)
    lo, hi = p["joint_limits"][:, 0], p["joint_limits"][:, 1]
    th = rng.uniform(lo, hi, (200, n))
    dth, ddth = rng.uniform(-2, 2, (200, n)), rng.uniform(-5, 5, (200, n))
    ft = rng.uniform(-10, 10, 6)
    g = np.array([0.3, -0.2, -9.81])
    for flags in (0, 1):
        rb = hostcheck.robot(p, flags)
        T, J = hostcheck.fk(rb, th)
        assert np.abs(T - o.forward_kinematics(th)).max() < 1e-12
        assert np.abs(J - o.jacobian(th)).max() < 1e-12
        ref = o.inverse_dynamics(th, dth, ddth, g, ft, analytic=True)
        got = hostcheck.rnea(rb, th, dth, ddth, g, ft)
        assert np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref).max(1, keepdims=True))) < 1e-11
        # the shared-memory state store used by the kernels gives the same bits as the register store
        assert np.array_equal(got, hostcheck.rnea(rb, th, dth, ddth, g, ft, smem_store=1))
        Mref = o.mass_matrix(th[:50])
        assert np.abs(hostcheck.mass(rb, th[:50]) - Mref).max() < 1e-9 * max(1, np.abs(Mref).max())
        tau = rng.uniform(-20, 20, (50, n))
        ref = o.forward_dynamics(th[:50], dth[:50], tau, g, ft, analytic=True)
        got = hostcheck.fd(rb, th[:50], dth[:50], tau, g, np.tile(ft, (50, 1)))
        assert np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref).max(1, keepdims=True))) < 1e-9


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8])
def test_general_inertia_and_prismatic_vs_oracle(hostcheck, n):
    """Arbitrary unit screws, a non-unit prismatic joint, full symmetric 6x6 inertias."""
    from oracle import Oracle

    p = random_general_pack(n, seed=100 + n)
    o = Oracle(p["S_list"], p["M"], p["Glist"], p["Mlist_per_link"])
    rb = hostcheck.robot(p)
    assert hostcheck.L.mpk_robot_is_rigid(rb[0]) == 0
    rng = np.random.default_rng(n)
    th, dth, ddth = rng.uniform(-2, 2, (40, n)), rng.uniform(-2, 2, (40, n)), rng.uniform(-3, 3, (40, n))
    g, ft = np.array([1.0, 2.0, -9.0]), rng.uniform(-5, 5, 6)
    T, J = hostcheck.fk(rb, th)
    assert np.abs(T - o.forward_kinematics(th)).max() < 1e-12
    assert np.abs(J - o.jacobian(th)).max() < 1e-12
    assert np.array_equal(hostcheck.rnea(rb, th, dth, ddth, g, ft), hostcheck.rnea(rb, th, dth, ddth, g, ft, smem_store=1))
    for analytic, tol in ((True, 1e-11), (False, 1e-7)):  # literal path carries its finite-difference noise
        ref = o.inverse_dynamics(th, dth, ddth, g, ft, analytic=analytic)
        got = hostcheck.rnea(rb, th, dth, ddth, g, ft)
        assert np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref).max(1, keepdims=True))) < tol
    Mref = o.mass_matrix(th)  # literal sum_k Jk^T Gk Jk
    assert np.abs(hostcheck.mass(rb, th) - Mref).max() < 1e-10 * max(1, np.abs(Mref).max())
    gref = o.gravity_forces(th, g)
    assert np.abs(hostcheck.rnea(rb, th, g=g) - gref).max() < 1e-10 * max(1, np.abs(gref).max())


def _awkward_pack(kind, seed):
    """Chains that stress the link-frame construction of csrc/robot.cu: consecutive axes that
    are exactly parallel, coincident, nearly parallel (the Hayati branch, beta != 0), or
    prismatic at either end."""
    from conftest import random_general_pack

    rng = np.random.default_rng(seed)
    n = 5
    p = random_general_pack(n, seed)
    S = np.zeros((6, n))

    def rev(w, q):
        w = np.asarray(w, float) / np.linalg.norm(w)
        return np.r_[w, -np.cross(w, q)]

    def tilt(w, eps):
        t = np.cross(w, rng.normal(size=3))
        return w + eps * t / np.linalg.norm(t)

    w0 = rng.normal(size=3)
    w0 /= np.linalg.norm(w0)
    q = lambda: rng.uniform(-0.5, 0.5, 3)
    if kind == "parallel":
        for i in range(n):
            S[:, i] = rev(w0 if i % 2 == 0 else -w0, q())
    elif kind == "coincident":
        q0 = q()
        S[:, 0] = rev(w0, q0)
        S[:, 1] = rev(w0, q0 + 0.3 * w0)          # same line
        S[:, 2] = rev(-w0, q0)                     # same line, opposite sense
        S[:, 3] = rev(rng.normal(size=3), q0)      # intersecting it
        S[:, 4] = rev(rng.normal(size=3), q())
    elif kind.startswith("tilt"):
        eps = float(kind[4:])
        w = w0
        for i in range(n):
            S[:, i] = rev(w, q())
            w = tilt(w / np.linalg.norm(w), eps)
    elif kind == "prismatic_ends":
        S[3:, 0] = w0 * 1.3
        for i in range(1, n - 1):
            S[:, i] = rev(rng.normal(size=3), q())
        S[3:, n - 1] = S[:3, n - 2] * 0.8          # sliding along the previous axis
    elif kind == "all_prismatic":
        for i in range(n):
            v = rng.normal(size=3)
            S[3:, i] = v / np.linalg.norm(v) * (0.5 + i)
    p["S_list"] = S
    return p


@pytest.mark.parametrize("rigid", [False, True], ids=["general", "rigid"])
@pytest.mark.parametrize("kind", ["parallel", "coincident", "tilt1e-2", "tilt1e-5", "tilt1e-9", "tilt0.099",
                                  "prismatic_ends", "all_prismatic"])
def test_awkward_axis_arrangements_vs_oracle(hostcheck, kind, rigid):
    from oracle import Oracle

    p = _awkward_pack(kind, seed=7)
    n = p["S_list"].shape[1]
    if rigid:  # rigid bodies at their centres of mass: block-diagonal inertias
        rng = np.random.default_rng(1)
        G = np.zeros((n, 6, 6))
        for i in range(n):
            A = rng.normal(size=(3, 3))
            G[i, :3, :3] = A @ A.T + 0.1 * np.eye(3)
            G[i, 3:, 3:] = np.eye(3) * rng.uniform(0.5, 3.0)
        p["Glist"] = G
    o = Oracle(p["S_list"], p["M"], p["Glist"], p["Mlist_per_link"])
    rb = hostcheck.robot(p)
    assert hostcheck.L.mpk_robot_is_rigid(rb[0]) == int(rigid)
    assert hostcheck.L.mpk_robot_all_revolute(rb[0]) == int("prismatic" not in kind)
    rng = np.random.default_rng(3)
    th, dth, ddth = rng.uniform(-2, 2, (30, n)), rng.uniform(-2, 2, (30, n)), rng.uniform(-3, 3, (30, n))
    g, ft = np.array([1.0, 2.0, -9.0]), rng.uniform(-5, 5, 6)
    T, J = hostcheck.fk(rb, th)
    assert np.abs(T - o.forward_kinematics(th)).max() < 1e-12
    assert np.abs(J - o.jacobian(th)).max() < 1e-12
    ref = o.inverse_dynamics(th, dth, ddth, g, ft, analytic=True)
    got = hostcheck.rnea(rb, th, dth, ddth, g, ft)
    assert np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref).max(1, keepdims=True))) < 1e-11
    assert np.array_equal(got, hostcheck.rnea(rb, th, dth, ddth, g, ft, smem_store=1))
    Mref = o.mass_matrix(th)
    assert np.abs(hostcheck.mass(rb, th) - Mref).max() < 1e-10 * max(1, np.abs(Mref).max())
    tau = rng.uniform(-20, 20, (30, n))
    ref = o.forward_dynamics(th, dth, tau, g, ft, analytic=True)
    got = hostcheck.fd(rb, th, dth, tau, g, np.tile(ft, (30, 1)))
    assert np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref).max(1, keepdims=True))) < 1e-8


def test_joint_sincos_accuracy(hostcheck):
    """The kernels' own sin/cos (Cody-Waite + fdlibm kernels with constant-bank coefficients,
    csrc/mpk_device.cuh sincos_pack) against libm: <= 2 ulp, exact
    symmetries at the quadrant boundaries, library fallback for huge / non-finite arguments."""
    rng = np.random.default_rng(0)
    k = np.arange(-4000, 4001)
    x = np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-1e5, 1e5, 100000), rng.normal(0, 1e-3, 1000),
                        k * (np.pi / 4), np.nextafter(k * (np.pi / 4), np.inf), np.nextafter(k * (np.pi / 2), -np.inf),
                        [0.0, -0.0, 1e-300, 5e-324, 99999.99999, -99999.99999]])
    s, c = hostcheck.sincos(x)
    rs, rc = np.sin(x), np.cos(x)
    assert np.max(np.abs(s - rs)) < 2.3e-16 and np.max(np.abs(c - rc)) < 2.3e-16
    ulp = lambda got, ref: np.max(np.abs(got - ref) / np.maximum(np.spacing(np.abs(ref)), 1e-300) * (np.abs(ref) > 1e-3))
    assert ulp(s, rs) <= 2.0 and ulp(c, rc) <= 2.0  # the CUDA library's own bound for sin / cos
    assert np.abs(s * s + c * c - 1).max() < 5e-16
    s0, c0 = hostcheck.sincos([0.0])
    assert s0[0] == 0.0 and c0[0] == 1.0
    big = np.array([1e5, 1.234e7, -3.3e12, 1e300])
    s, c = hostcheck.sincos(big)
    assert np.array_equal(s, np.sin(big)) and np.array_equal(c, np.cos(big))
    s, c = hostcheck.sincos([np.nan, np.inf, -np.inf])
    assert np.isnan(s).all() and np.isnan(c).all()


@pytest.mark.parametrize("robot", ROBOTS)
def test_float32_kernel_algebra_within_north_star_tolerance(hostcheck, oracle_factory, robot):
    """The float32 instantiation of the kernel templates: torques within 1e-4 relative and poses
    within 1e-5 of the float64 oracle (BASELINE.json north_star: "fp32 kernels")."""
    p = load_pack(robot)
    o = oracle_factory(robot)
    n = p["S_list"].shape[1]
    rng = np.random.default_rng(5)
    lo, hi = p["joint_limits"][:, 0], p["joint_limits"][:, 1]
    th = rng.uniform(lo, hi, (300, n))
    dth, ddth = rng.uniform(-2, 2, (300, n)), rng.uniform(-5, 5, (300, n))
    g = np.array([0.0, 0.0, -9.81])
    for flags in (0, 1):
        rb = hostcheck.robot(p, flags)
        ref = o.inverse_dynamics(th, dth, ddth, g, None, analytic=True)
        got = hostcheck.rnea(rb, th, dth, ddth, g, None, f32=True)
        assert np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref).max(1, keepdims=True))) < 1e-4
        ft = rng.uniform(-10, 10, 6)
        ref = o.inverse_dynamics(th, dth, ddth, g, ft, analytic=True)
        got = hostcheck.rnea(rb, th, dth, ddth, g, ft, f32=True)
        assert np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref).max(1, keepdims=True))) < 1e-4
        T, J = hostcheck.fk(rb, th, f32=True)
        Tref, Jref = o.forward_kinematics(th), o.jacobian(th)
        assert np.abs(T - Tref).max() < 1e-5 * max(1, np.abs(Tref).max())
        assert np.abs(J - Jref).max() < 1e-5 * max(1, np.abs(Jref).max())


def test_planar_2r_known_answers(hostcheck):
    """Murray-Li-Sastry Ex. 4.3 (reference tests/test_v132_regressions.py:126-192, 229-286)."""
    rb = hostcheck.robot(planar_2r_pack())
    th = np.array([[0.0, np.pi / 2]])
    c2 = np.cos(th[0, 1])
    Mexp = np.array([[1 + (1 + 1 + 2 * c2), 1 + c2], [1 + c2, 1.0]])
    np.testing.assert_allclose(hostcheck.mass(rb, th)[0], Mexp, atol=1e-12)
    np.testing.assert_allclose(hostcheck.rnea(rb, th, g=[-9.81, 0, 0])[0], [-9.81, -9.81], atol=1e-12)
    np.testing.assert_allclose(hostcheck.rnea(rb, th, g=[0, -9.81, 0])[0], [19.62, 0.0], atol=1e-12)
    np.testing.assert_allclose(hostcheck.rnea(rb, th, g=[0, 0, -9.81])[0], [0.0, 0.0], atol=1e-12)


def test_time_scaling_bit_exact_vs_golden(hostcheck):
    """The kernel's time scaling + rounding restated on the host is bit-identical to the oracle
    (itself pinned bit-exact on the reference in test_oracle_golden.py)."""
    from oracle import Oracle

    g = load_golden("trajectory")
    lim = g["joint_limits"]
    for name in ("cfg1", "cubic50", "two", "clipped", "odd_tf"):
        Tf, N, method = g[f"{name}_args"]
        got = hostcheck.traj(g[f"{name}_start"], g[f"{name}_end"], Tf, int(N), int(method), lim)
        ref = Oracle.joint_trajectory(g[f"{name}_start"], g[f"{name}_end"], Tf, int(N), int(method), lim)
        for a, k in zip(got, ("positions", "velocities", "accelerations")):
            assert np.array_equal(a.view(np.uint32), ref[k].view(np.uint32)), (name, k)
    # other methods -> zero scaling; N = 1 -> NaN (the planner's CPU contract)
    got = hostcheck.traj(np.zeros(3), np.ones(3), 2.0, 5, 7, None)
    assert np.all(got[0] == 0) and np.all(got[1] == 0)
    got = hostcheck.traj(np.zeros(3), np.ones(3), 2.0, 1, 5, None)
    ref = Oracle.joint_trajectory(np.zeros(3), np.ones(3), 2.0, 1, 5)
    assert np.array_equal(np.isnan(got[0]), np.isnan(ref["positions"]))


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_inverse_kinematics_kernel_vs_reference_golden(hostcheck, robot):
    """The kernel's per-target DLS solver (csrc/mpk_device.cuh ik_dls) against the unmodified
    reference: same success flags and iteration counts, solutions to 1e-7, for every run that
    does not reach the stagnation restart (whose noise comes from a different generator);
    exhausted runs report max_iterations + 1 and failure like the reference."""
    g = load_golden("inverse_kinematics")
    rb = hostcheck.kin_robot(g[f"{robot}_S"], g[f"{robot}_M"])
    lim = g[f"{robot}_limits"]
    for i, (Td, seed, par) in enumerate(zip(g[f"{robot}_T"], g[f"{robot}_seed"], g[f"{robot}_params"])):
        th, ok, it = hostcheck.ik(rb, Td, seed, max_iterations=int(par[0]), damping=par[1], step_cap=par[2],
                                  weight_orientation=par[3], weight_position=par[4], limits=lim)
        assert bool(ok[0]) == bool(g[f"{robot}_success"][i]), i
        if g[f"{robot}_success"][i]:
            assert int(it[0]) == int(g[f"{robot}_iterations"][i]), i
            np.testing.assert_allclose(th[0], g[f"{robot}_theta"][i], rtol=0, atol=1e-7, err_msg=str(i))
        else:
            assert int(it[0]) == int(par[0]) + 1
    # with the restart noise taken from NumPy's global generator (the Python mirror's protocol for
    # single-target calls) EVERY golden run is reproduced, the ones through stagnation restarts too,
    # and the generator ends where the reference leaves it
    for i, (Td, seed, par) in enumerate(zip(g[f"{robot}_T"], g[f"{robot}_seed"], g[f"{robot}_params"])):
        np.random.seed(100 + i)
        th, ok, it = hostcheck.ik(rb, Td, seed, max_iterations=int(par[0]), damping=par[1], step_cap=par[2],
                                  weight_orientation=par[3], weight_position=par[4], limits=lim, numpy_restarts=True)
        assert bool(ok[0]) == bool(g[f"{robot}_success"][i]) and int(it[0]) == int(g[f"{robot}_iterations"][i]), i
        np.testing.assert_allclose(th[0], g[f"{robot}_theta"][i], rtol=0, atol=1e-7, err_msg=str(i))
    # batch call = per-target calls; zero iterations allowed
    Tds, seeds = g[f"{robot}_T"][[0, 3, 4]], g[f"{robot}_seed"][[0, 3, 4]]
    th, ok, it = hostcheck.ik(rb, Tds, seeds, max_iterations=300, limits=lim)
    assert ok.all() and np.array_equal(it, g[f"{robot}_iterations"][[0, 3, 4]])
    th0, ok0, it0 = hostcheck.ik(rb, Tds, seeds, max_iterations=0, limits=lim)
    assert not ok0.any() and np.array_equal(th0, seeds) and (it0 == 1).all()


@pytest.mark.parametrize("mode", ["adaptive", "backtracking", "both"])
@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_inverse_kinematics_kernel_modes_vs_reference_golden(hostcheck, robot, mode):
    """... with adaptive_tuning and / or backtracking (kinematics/ik.py:215-229, 253-276): same
    success flags and iteration counts as the unmodified reference, solutions to 1e-7."""
    g = load_golden("inverse_kinematics_modes")
    rb = hostcheck.kin_robot(g[f"{robot}_S"], g[f"{robot}_M"])
    kw = dict(adaptive_tuning=mode in ("adaptive", "both"), backtracking=mode in ("backtracking", "both"))
    for i, (Td, seed, budget) in enumerate(zip(g[f"{robot}_T"], g[f"{robot}_seed"], g[f"{robot}_max_iterations"])):
        th, ok, it = hostcheck.ik(rb, Td, seed, max_iterations=int(budget), limits=g[f"{robot}_limits"], **kw)
        assert bool(ok[0]) == bool(g[f"{robot}_{mode}_success"][i]), i
        if ok[0]:
            assert int(it[0]) == int(g[f"{robot}_{mode}_iterations"][i]), i
            np.testing.assert_allclose(th[0], g[f"{robot}_{mode}_theta"][i], rtol=0, atol=1e-7, err_msg=str(i))
        else:
            assert int(it[0]) == int(budget) + 1


def test_cartesian_trajectory_kernel_vs_reference_golden(hostcheck):
    """The kernel's per-step Cartesian interpolation (csrc/mpk_device.cuh cartesian_point) against
    the unmodified reference, to float32 rounding."""
    g = load_golden("cartesian_trajectory")
    for name in ("generic5", "generic3", "method1", "same_R", "tiny", "near_pi", "pi_band", "pi_exact"):
        Tf, N, method = g[f"{name}_args"]
        got = hostcheck.cartesian(g[f"{name}_Xstart"], g[f"{name}_Xend"], float(Tf), int(N), int(method))
        for k in ("positions", "velocities", "accelerations", "orientations"):
            np.testing.assert_allclose(got[k], g[f"{name}_{k}"], rtol=3e-7, atol=1e-7, err_msg=f"{name} {k}")


def test_body_frame_kinematics_vs_reference_golden(hostcheck):
    """frame="body": the kernel template with screws S' = Ad(M) B and the Ad(T^-1) column
    transform against the unmodified reference (tests/golden/body_kinematics.npz)."""
    from manipulapy_b200.kinematics import _adjoint

    g = load_golden("body_kinematics")
    for k in ("ur5", "free"):
        M, B, th = g[f"{k}_M"], g[f"{k}_B"], g[f"{k}_theta"]
        rb = hostcheck.kin_robot(_adjoint(M) @ B, M)
        T, J = hostcheck.fk(rb, th, body=True)
        np.testing.assert_allclose(T, g[f"{k}_T"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(J, g[f"{k}_J"], rtol=0, atol=1e-12)


def test_registry_trajectory_contract_vs_golden(hostcheck):
    """The kernel's time scaling under MPK_TRAJ_REGISTRY_CONTRACT against the reference's registry
    launcher outputs: linear for other methods, sit-at-start for N <= 1 / Tf <= 0.  The reference
    computes this path in float32 (and its CUDA kernels with fast-math), the kernel in float64
    with one rounding: equal to float32 rounding noise."""
    g = load_golden("registry_trajectory")
    for name in ("linear", "method7", "cubic", "quintic", "n1", "tf0", "tfneg"):
        Tf, N, method = g[f"{name}_args"]
        got = hostcheck.traj(g[f"{name}_start"], g[f"{name}_end"], float(Tf), int(N), int(method) | 0x100, None)
        for a, k in zip(got, ("positions", "velocities", "accelerations")):
            ref = g[f"{name}_{k}"]
            # (the reference's float32 polynomial carries a few float32 ulps of its largest term)
            np.testing.assert_allclose(a, ref, rtol=2e-6, atol=4e-6 * max(1.0, float(np.abs(ref).max(initial=0.0))),
                                       err_msg=f"{name} {k}")
    # the planner contract is untouched: other methods -> zero scaling
    got = hostcheck.traj(np.zeros(3), np.ones(3), 2.0, 5, 1, None)
    assert np.all(got[0] == 0) and np.all(got[1] == 0)


def test_public_signatures_match_reference_api_contract():
    """tests/data/api_contract_golden.json of the reference, hot-path subset: every mirrored
    method takes the reference's parameters, same names, same order; parameters that are optional
    in the reference are optional here (extension keywords may follow)."""
    import inspect
    import json

    import manipulapy_b200 as mp

    contract = json.loads((GOLDEN_DIR / "api_contract.json").read_text())
    for key, spec in contract.items():
        if key.startswith("_"):
            continue
        cls, meth = key.split(".")
        params = [p for p in inspect.signature(getattr(getattr(mp, cls), meth)).parameters.values()
                  if p.name != "self"]
        want = spec["parameters"]
        assert [p.name for p in params[:len(want)]] == [w["name"] for w in want], key
        for p, w in zip(params, want):
            assert p.kind in (p.POSITIONAL_OR_KEYWORD,), (key, p.name)
            if w["has_default"]:
                assert p.default is not inspect.Parameter.empty, (key, p.name)
        for p in params[len(want):]:  # extensions never become required
            assert p.default is not inspect.Parameter.empty or p.kind == p.KEYWORD_ONLY, (key, p.name)


def test_registry_contract():
    from manipulapy_b200 import KERNEL_REGISTRY, KernelRegistration

    names = KERNEL_REGISTRY.names()
    for v in ("auto", "auto_tune", "standard", "vectorized", "memory_optimized", "warp_optimized", "cache_friendly"):
        assert f"trajectory.{v}" in names
    assert "dynamics.inverse" in names and "dynamics.forward_rollout" in names
    with pytest.raises(KeyError, match="Unknown CUDA kernel 'nope'. Available kernels:"):
        KERNEL_REGISTRY.get("nope")
    e = KERNEL_REGISTRY.get("trajectory.auto")
    with pytest.raises(ValueError, match="already registered"):
        KERNEL_REGISTRY.register(e)
    with pytest.raises(TypeError):
        e.metadata["x"] = 1
    with pytest.raises(RuntimeError, match="no CPU"):
        e.cpu_launcher()
    assert isinstance(e, KernelRegistration)


def test_no_cpu_fallback():
    """Without CUDA the product path refuses to run; it never routes through the oracle."""
    import torch

    from manipulapy_b200 import load_robot

    rb = load_robot("ur5")
    with pytest.raises(RuntimeError, match="no CPU"):
        rb.planner(use_cuda=False)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            rb.planner()
        with pytest.raises(RuntimeError):
            rb.dynamics.mass_matrix(np.zeros(6))
    src = "".join(p.read_text() for p in (REPO / "manipulapy_b200").rglob("*.py"))
    assert "import oracle" not in src and "from oracle" not in src


def test_legacy_path_raises():
    from manipulapy_b200 import ManipulatorDynamics

    p = load_pack("ur5")
    with pytest.raises(NotImplementedError, match="legacy"):
        ManipulatorDynamics(p["M"], None, None, None, p["S_list"], None, p["Glist"], None)


def test_shard_range():
    from manipulapy_b200 import shard_range

    for units in (0, 1, 7, 4096, 9998336):
        for world in (1, 2, 4, 8):
            spans = [shard_range(units, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == units
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == -(-units // world)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from manipulapy_b200.sharding import shard_range, gather_rows
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2],
                        rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
units = 11
full = torch.arange(units * 3, dtype=torch.float32).reshape(units, 3)
lo, hi = shard_range(units, 2, rank)
out = gather_rows(full[lo:hi].clone(), units)
assert torch.equal(out, full), out
out = gather_rows(full[lo:hi].clone(), units, dst=0)
assert (out is None) if rank else torch.equal(out, full)
# chunked gather that overlaps the transfer of chunk k with the computation of chunk k + 1
from manipulapy_b200.sharding import gather_rows_pipelined
calls = []
def launch(a, b, dest):
    calls.append((a, b))
    dest.copy_(full[a:b] * 2)
for units2, chunks in ((11, 3), (11, 1), (2, 8), (1, 4), (0, 2)):
    f2 = full[:units2]
    calls.clear()
    out = gather_rows_pipelined(launch, units2, (3,), torch.float32, "cpu", dst=0, chunks=chunks)
    lo2, hi2 = shard_range(units2, 2, rank)
    assert sorted(calls) == calls and sum(b - a for a, b in calls) == hi2 - lo2, calls
    assert (out is None) if rank else torch.equal(out, f2 * 2), (units2, chunks, out)
dist.destroy_process_group()
print("ok", rank)
"""


def test_shard_bounds_weighted():
    from manipulapy_b200.sharding import shard_bounds, shard_range

    assert shard_bounds(10, 4) == [0, 3, 6, 9, 10]
    assert shard_bounds(0, 3) == [0, 0, 0, 0]
    assert shard_bounds(2, 4) == [0, 1, 2, 2, 2]
    # device->host rates of an 8-GPU box whose GPUs 0-3 share slower uplinks (profiles/r1_d2h_n8.json)
    w = [11.7] * 4 + [18.5] * 4
    b = shard_bounds(32768, 8, w)
    assert b[0] == 0 and b[-1] == 32768 and all(y >= x for x, y in zip(b, b[1:]))
    sizes = [y - x for x, y in zip(b, b[1:])]
    assert abs(sizes[0] / sizes[7] - 11.7 / 18.5) < 0.01 and sum(sizes) == 32768
    assert shard_range(32768, 8, 3, w) == (b[3], b[4])
    for bad in ([1.0] * 7, [1.0] * 7 + [0.0], [1.0] * 7 + [float("nan")]):
        with pytest.raises(ValueError):
            shard_bounds(10, 8, bad)
    # aligned interior boundaries (a shard of a shared float32 result buffer whose rows are 8 mod 16 bytes long must
    # start on an even row to be 16-byte aligned): still a partition, every interior boundary a multiple of `align`
    for units, world, wts in ((4097, 8, None), (410, 8, None), (37, 2, None), (37, 2, [1.0, 3.0]), (1, 4, None), (0, 3, None),
                              (32768, 8, w)):
        for align in (1, 2, 4):
            bb = shard_bounds(units, world, wts, align)
            assert bb[0] == 0 and bb[-1] == units and len(bb) == world + 1
            assert all(y >= x for x, y in zip(bb, bb[1:]))
            assert all(x % align == 0 or x == units for x in bb[1:-1])
            assert shard_range(units, world, world - 1, wts, align) == (bb[-2], bb[-1])
    assert shard_bounds(4097, 8, None, 2) == [0, 514, 1028, 1542, 2056, 2570, 3084, 3598, 4097]


def test_gather_rows_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(REPO), port, str(r)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


# ---------------------------------------------------------------------------------------------
# every robot of the reference's bundled database (oracle/gen_robot_zoo.py): the kernels' link
# frame construction (Denavit-Hartenberg / Hayati) on 26 different axis arrangements
# ---------------------------------------------------------------------------------------------
from conftest import ZOO_ROBOTS, check_zoo_outputs, load_zoo  # noqa: E402


@pytest.mark.parametrize("robot", ZOO_ROBOTS)
def test_robot_zoo_kernel_algebra_vs_reference(hostcheck, robot):
    z = load_zoo()[robot]
    th, dth, ddth, ft, g = z["thetas"], z["dthetas"], z["ddthetas"], z["ftips"], z["g"]
    for flags in (0, 1):
        rb = hostcheck.robot(z, flags)
        T, J = hostcheck.fk(rb, th)
        check_zoo_outputs(
            z, T, J, hostcheck.mass(rb, th), hostcheck.rnea(rb, th, g=g),
            hostcheck.rnea(rb, th, dth, g=(0, 0, 0)), hostcheck.rnea(rb, th, dth, ddth, g, ft),
            hostcheck.fd(rb, th, dth, z["taus"], g, ft))


def _front_end_cases(hostcheck, robot):
    g = load_golden("inverse_kinematics_front_ends")
    rb = hostcheck.kin_robot(g[f"{robot}_S"], g[f"{robot}_M"])
    lim = g[f"{robot}_limits"]
    limits = [tuple(r) for r in lim]
    n = lim.shape[0]
    fk = lambda th: hostcheck.fk(rb, th)[0]  # noqa: E731
    return g, rb, lim, limits, n, fk


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_smart_inverse_kinematics_kernel_vs_reference_golden(hostcheck, robot):
    """smart_inverse_kinematics = ik_helpers.smart_driver around the kernel's solver, against the
    unmodified reference: success flags, total iteration counts and solutions of EVERY golden case
    (single-target protocol: restart noise from NumPy's generator); the batch driver (counter-based
    restart noise) solves what the per-target calls solve."""
    from manipulapy_b200 import ik_helpers

    g, rb, lim, limits, n, fk = _front_end_cases(hostcheck, robot)

    def solve(Tds, th0, numpy_restarts=False):
        return hostcheck.ik(rb, Tds, th0, max_iterations=120, limits=lim, adaptive_tuning=True, backtracking=True,
                            numpy_restarts=numpy_restarts)

    # one target per call, restart noise from NumPy's generator (what SerialManipulator.smart_inverse_kinematics
    # does for a single target): every golden case, fall-back starts and restarts included
    for i, Td in enumerate(g[f"{robot}_T"]):
        np.random.seed(300 + i)
        th, ok, it = ik_helpers.smart_driver(lambda T, t: solve(T, t, True), fk, Td[None], n, limits,
                                             "workspace_heuristic", True)
        assert bool(ok[0]) == bool(g[f"{robot}_smart_success"][i]), i
        assert int(it[0]) == int(g[f"{robot}_smart_iterations"][i]), i
        np.testing.assert_allclose(th[0], g[f"{robot}_smart_theta"][i], rtol=0, atol=1e-7, err_msg=str(i))
    # the batch driver: every target the per-target calls solve without restarts is solved
    np.random.seed(1)
    th, ok, it = ik_helpers.smart_driver(solve, fk, g[f"{robot}_T"], n, limits, "workspace_heuristic", True)
    easy = g[f"{robot}_smart_success"] & (g[f"{robot}_smart_restarts"] == 0)
    assert ok[easy].all() and not ok[-1]
    assert ik_helpers.pose_error(fk(th[ok]), g[f"{robot}_T"][ok]).max() < 1e-5


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_robust_inverse_kinematics_kernel_vs_reference_golden(hostcheck, robot):
    """robust_inverse_kinematics = ik_helpers.robust_driver around the kernel's solver: success, total
    iterations, winning strategy and solution as the unmodified reference, every golden case."""
    from manipulapy_b200 import ik_helpers

    g, rb, lim, limits, n, fk = _front_end_cases(hostcheck, robot)

    def solve(Tds, th0, damping, step_cap):
        return hostcheck.ik(rb, Tds, th0, eomg=2e-3, ev=2e-3, max_iterations=120, damping=damping,
                            step_cap=step_cap, limits=lim, adaptive_tuning=True, backtracking=True, numpy_restarts=True)

    for i, Td in enumerate(g[f"{robot}_T"]):
        np.random.seed(400 + i)
        th, ok, it, win = ik_helpers.robust_driver(solve, fk, Td[None], n, limits, 4)
        assert bool(ok[0]) == bool(g[f"{robot}_robust_success"][i]), i
        assert int(it[0]) == int(g[f"{robot}_robust_iterations"][i]), i
        assert str(win[0]) == str(g[f"{robot}_robust_strategy"][i]), i
        np.testing.assert_allclose(th[0], g[f"{robot}_robust_theta"][i], rtol=0, atol=1e-6, err_msg=str(i))


def test_ik_guess_helpers_match_reference_formulas():
    """The vectorised random guesses consume NumPy's global generator exactly like the reference's
    joint-by-joint draws (kinematics/ik_helpers.py:179-212); midpoint and clip as the reference."""
    from manipulapy_b200 import ik_helpers

    lim = [(-1.0, 2.0), (-3.0, 0.5), (0.1, 0.2), (-np.pi, np.pi)]
    np.random.seed(5)
    a = ik_helpers.random_in_limits_batch(lim, 7)
    np.random.seed(5)
    b = np.stack([ik_helpers.random_in_limits(lim) for _ in range(7)])
    assert np.array_equal(a, b)
    assert np.random.uniform() == np.random.RandomState(5).uniform(size=7 * 4 + 1)[-1]  # same stream position
    open_lim = [(-1.0, None), (None, 2.0), (None, None), (0.0, 1.0)]
    np.random.seed(6)
    c = ik_helpers.random_in_limits_batch(open_lim, 3)
    np.random.seed(6)
    d = np.stack([ik_helpers.random_in_limits(open_lim) for _ in range(3)])
    assert np.array_equal(c, d) and (c[:, 0] >= -1).all() and (c[:, 1] <= 2).all()
    assert np.array_equal(ik_helpers.midpoint_of_limits(open_lim), [0.0, 0.0, 0.0, 0.5])
    assert ik_helpers.random_in_limits_batch(lim, 0).shape == (0, 4)
    T = np.eye(4)
    T[:3, 3] = [0.3, 0.2, 0.4]
    g1 = ik_helpers.workspace_heuristic_guess(T, 6, [(-3, 3)] * 6)
    gb = ik_helpers.workspace_heuristic_guess(np.stack([T, T]), 6, [(-3, 3)] * 6)
    assert g1.shape == (6,) and np.array_equal(gb[0], g1) and np.isclose(g1[0], np.arctan2(0.2, 0.3))


# ---------------------------------------------------------------------------------------------
# link-geometry signatures: the kernels compiled for an arm family's parallel / perpendicular /
# intersecting axes return what the general kernels return
# ---------------------------------------------------------------------------------------------
def test_link_geometry_signatures_of_the_robot_database(hostcheck):
    zoo = load_zoo()
    sig = {r: hostcheck.geo(hostcheck.robot(zoo[r])) for r in ZOO_ROBOTS}
    assert sig["ur5"] == sig["ur10e"] == sig["ur3"] == 0xD52AD0   # axes 2, 3, 4 parallel; 1, 5, 6 perpendicular
    assert sig["iiwa14"] == sig["iiwa7"] == 0xD555150             # all perpendicular, all but one intersecting
    assert sig["crx10ia"] == 0xD558D0 and sig["abb_irb2400"] == 0xDD1A90 and sig["fanuc_lrmate"] == 0xDD1890
    assert sig["panda"] == 0                                      # not a plain revolute chain: general kernels
    assert sig["xarm6"] == 0x8402C0                               # alpha = pi/2 only to 1e-6: class bits a = 0 / d = 0 only


@pytest.mark.parametrize("robot", ["ur5", "ur10e", "iiwa14", "crx10ia", "fanuc_lrmate", "abb_irb2400", "gen3"])
def test_geometry_specialised_kernels_equal_general_kernels(hostcheck, robot):
    z = load_zoo()[robot]
    rb = hostcheck.robot(z)
    geo = hostcheck.geo(rb)
    assert geo != 0
    rng = np.random.default_rng(8)
    n = rb[1]
    th, dth, ddth = rng.uniform(-3, 3, (64, n)), rng.uniform(-2, 2, (64, n)), rng.uniform(-5, 5, (64, n))
    tau, ft = rng.uniform(-10, 10, (64, n)), rng.uniform(-5, 5, 6)

    def run():
        return (hostcheck.rnea(rb, th, dth, ddth), hostcheck.rnea(rb, th, dth, ddth, smem_store=True),
                hostcheck.rnea(rb, th, dth, ddth, ftip=ft), hostcheck.rnea(rb, th), hostcheck.mass(rb, th),
                hostcheck.fd(rb, th, dth, tau))

    special = run()
    assert hostcheck.set_geo(rb, 0) == geo
    general = run()
    hostcheck.set_geo(rb, geo)
    for a, b in zip(special, general):
        sc = np.maximum(1.0, np.abs(b).reshape(64, -1).max(1)).reshape((64,) + (1,) * (b.ndim - 1))
        assert (np.abs(a - b) / sc).max() < 5e-14
