#!/usr/bin/env python3
"""Generate golden fixtures by running the UNMODIFIED reference (ManipulaPy v1.4.1).

TEST INFRASTRUCTURE ONLY.  Runs in the build container, where the reference is
mounted read-only at /root/reference (it does not exist on the GPU box, so the
vectors are committed):

    python oracle/gen_golden.py

Writes
  manipulapy_b200/robots/<robot>.npz   constant packs (S_list, M, Glist, Mlist_per_link,
                                       joint_limits) extracted by the reference's own
                                       URDFToSerialManipulator -- shipped as robot data.
  tests/golden/dynamics_<robot>.npz    per-point float64 vectors: the reference's own
                                       golden file replayed verbatim (ur5, panda:
                                       tests/data/dynamics_golden_*.npz) plus FK / Jacobian /
                                       forward_dynamics outputs computed here by the reference.
  tests/golden/trajectory.npz          joint_trajectory / batch_joint_trajectory (float32).
  tests/golden/body_kinematics.npz     forward_kinematics / jacobian with frame="body".
  tests/golden/cartesian_trajectory.npz cartesian_trajectory (positions, velocities, accelerations, orientations).
  tests/golden/inverse_kinematics.npz  iterative_inverse_kinematics (theta, success, iterations).
  tests/golden/registry_trajectory.npz the registry launcher seam (linear method, N <= 1 / Tf <= 0 guards).
  tests/golden/id_trajectory.npz       inverse_dynamics_trajectory (float32, clipped).
  tests/golden/fd_trajectory.npz       forward_dynamics_trajectory rollouts (float32).

The reference needs a 2-file matplotlib stub (its only missing hard import,
planning/_kernels.py:27); the stub is created in a temp dir, not in the repo.
"""

from __future__ import annotations

import logging
import os
import sys
import tempfile
import warnings
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
REF = Path(os.environ.get("MANIPULAPY_REFERENCE", "/root/reference"))


def _bootstrap_reference() -> None:
    stub = Path(tempfile.mkdtemp(prefix="mpl_stub_")) / "matplotlib"
    stub.mkdir(parents=True)
    (stub / "__init__.py").write_text("def use(*a, **k):\n    pass\n")
    (stub / "pyplot.py").write_text("def __getattr__(name):\n    raise AttributeError(name)\n")
    sys.path[:0] = [str(REF), str(stub.parent)]
    os.environ.setdefault("MANIPULAPY_QUIET", "1")
    os.environ.setdefault("NUMBA_DISABLE_CUDA", "1")
    os.environ.setdefault("NUMBA_CACHE_DIR", tempfile.mkdtemp(prefix="numba_cache_"))
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    sys.dont_write_bytecode = True
    warnings.filterwarnings("ignore")
    logging.disable(logging.CRITICAL)


_bootstrap_reference()

import numpy as np  # noqa: E402

from ManipulaPy.ManipulaPy_data import get_robot_urdf  # noqa: E402
from ManipulaPy.path_planning import OptimizedTrajectoryPlanning  # noqa: E402
from ManipulaPy.urdf_processor import URDFToSerialManipulator  # noqa: E402

G_VEC = np.array([0.0, 0.0, -9.81])
ROBOT_DIR = REPO / "manipulapy_b200" / "robots"
GOLD_DIR = REPO / "tests" / "golden"


def load(robot: str):
    proc = URDFToSerialManipulator(get_robot_urdf(robot), load_meshes=False)
    return proc, proc.serial_manipulator, proc.dynamics


def limits_array(proc, n: int) -> np.ndarray:
    lims = np.empty((n, 2))
    for i, (lo, hi) in enumerate(proc.robot_data["joint_limits"]):
        lims[i] = (-np.pi if lo is None else lo, np.pi if hi is None else hi)
    return lims


def clear_caches(dyn) -> None:
    dyn._mass_matrix_cache.clear()
    dyn._mass_matrix_derivative_cache.clear()


def export_robot(robot: str) -> None:
    proc, sm, dyn = load(robot)
    n = dyn.S_list.shape[1]
    np.savez(
        ROBOT_DIR / f"{robot}.npz",
        S_list=np.asarray(dyn.S_list, np.float64),
        M=np.asarray(dyn.M_list, np.float64),
        Glist=np.asarray(dyn.Glist, np.float64),
        Mlist_per_link=np.asarray(dyn.Mlist_per_link, np.float64),
        joint_limits=limits_array(proc, n),
    )
    print(f"robot pack {robot}: n={n}")


def dynamics_golden(robot: str, n_cfg: int = 12, n_fd: int = 6) -> None:
    proc, sm, dyn = load(robot)
    n = dyn.S_list.shape[1]
    ref_npz = REF / "tests" / "data" / f"dynamics_golden_{robot}.npz"
    out = {}
    if ref_npz.exists():
        # the reference's own golden vectors, replayed verbatim
        with np.load(ref_npz) as d:
            for k in d.files:
                out[k] = d[k]
        out["source"] = np.array("reference tests/data/dynamics_golden_%s.npz" % robot)
    else:
        rng = np.random.default_rng(20260705)
        lims = limits_array(proc, n)
        th = rng.uniform(lims[:, 0], lims[:, 1], size=(n_cfg, n))
        th[0] = 0.0
        dth = rng.uniform(-1, 1, size=(n_cfg, n))
        ddth = rng.uniform(-1, 1, size=(n_cfg, n))
        dth[0] = ddth[0] = 0.0
        ft = np.zeros((n_cfg, 6))
        ft[min(5, n_cfg - 1)] = [1.0, -2.0, 0.5, 3.0, -1.5, 0.75]
        out.update(thetas=th, dthetas=dth, ddthetas=ddth, g=G_VEC.copy(), ftips=ft)
        M, ID, GF, C = [], [], [], []
        for i in range(n_cfg):
            clear_caches(dyn)
            M.append(np.asarray(dyn.mass_matrix(th[i])))
            ID.append(np.asarray(dyn.inverse_dynamics(th[i], dth[i], ddth[i], G_VEC, ft[i])))
            GF.append(np.asarray(dyn.gravity_forces(th[i], G_VEC)))
            C.append(np.asarray(dyn.velocity_quadratic_forces(th[i], dth[i])))
        out.update(mass_matrix=np.array(M), inverse_dynamics=np.array(ID),
                   gravity_forces=np.array(GF), velocity_quadratic_forces=np.array(C))
        out["source"] = np.array("reference run by oracle/gen_golden.py")
    th = out["thetas"]
    # FK / Jacobian have no golden in the reference: computed by the reference here.
    out["forward_kinematics"] = np.array([np.asarray(sm.forward_kinematics(t)) for t in th])
    out["jacobian"] = np.array([np.asarray(sm.jacobian(t)) for t in th])
    # forward_dynamics: a few rows (each call is 60-130 ms)
    rng = np.random.default_rng(7)
    idx = np.arange(min(n_fd, th.shape[0]))
    tau = rng.uniform(-20, 20, size=(idx.size, n))
    fd = []
    for r, i in enumerate(idx):
        clear_caches(dyn)
        fd.append(np.asarray(dyn.forward_dynamics(th[i], out["dthetas"][i], tau[r], G_VEC, out["ftips"][i])))
    out.update(fd_index=idx, fd_tau=tau, forward_dynamics=np.array(fd))
    np.savez(GOLD_DIR / f"dynamics_{robot}.npz", **out)
    print(f"dynamics golden {robot}: {th.shape[0]} configs, {idx.size} FD rows")


def make_planner(robot: str, torque_limits=None):
    proc, sm, dyn = load(robot)
    n = dyn.S_list.shape[1]
    lims = limits_array(proc, n)
    planner = OptimizedTrajectoryPlanning(
        sm, get_robot_urdf(robot), dyn, lims, torque_limits, use_cuda=False)
    return planner, dyn, lims


def trajectory_golden() -> None:
    planner, dyn, lims = make_planner("ur5")
    n = 6
    out = {"joint_limits": lims}
    rng = np.random.default_rng(1)
    cases = {
        # BASELINE config 1: seed 1, U(-1,1)^6, Tf=2, N=1000, quintic
        "cfg1": (rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), 2.0, 1000, 5),
        "cubic50": (rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), 1.5, 50, 3),
        "two": (rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), 0.7, 2, 5),
        # beyond the +-2pi / +-pi limits: exercises the position clip
        "clipped": (np.full(n, -7.0), np.full(n, 7.0), 3.0, 101, 5),
        "odd_tf": (rng.uniform(-3, 3, n), rng.uniform(-3, 3, n), 0.37, 257, 3),
    }
    for name, (s, e, Tf, N, method) in cases.items():
        r = planner.joint_trajectory(s, e, Tf, N, method)
        out[f"{name}_start"] = s
        out[f"{name}_end"] = e
        out[f"{name}_args"] = np.array([Tf, N, method], np.float64)
        for k in ("positions", "velocities", "accelerations"):
            a = np.asarray(r[k])
            assert a.dtype == np.float32
            out[f"{name}_{k}"] = a
    sb = rng.uniform(-3.5, 3.5, (5, n))
    eb = rng.uniform(-3.5, 3.5, (5, n))
    r = planner.batch_joint_trajectory(sb, eb, 2.0, 33, 5)
    out.update(batch_start=sb, batch_end=eb, batch_args=np.array([2.0, 33, 5], np.float64))
    for k in ("positions", "velocities", "accelerations"):
        out[f"batch_{k}"] = np.asarray(r[k])
    np.savez(GOLD_DIR / "trajectory.npz", **out)
    print("trajectory golden written")


def cartesian_golden() -> None:
    """cartesian_trajectory of the unmodified reference: a generic pose pair (quintic, cubic and
    a method that is neither), equal orientations, a tiny rotation, and rotations just below
    and inside the half-turn band of MatrixLog3."""
    planner, dyn, lims = make_planner("ur5")
    rng = np.random.default_rng(8)

    def rot(axis, ang):
        a = np.asarray(axis, float) / np.linalg.norm(axis)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K

    def pose(R, p):
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = p
        return T

    R0 = rot(rng.normal(size=3), 0.7)
    ax = rng.normal(size=3)
    cases = {
        "generic5": (pose(R0, rng.uniform(-1, 1, 3)), pose(rot(rng.normal(size=3), 1.9), rng.uniform(-1, 1, 3)), 2.0, 40, 5),
        "generic3": (pose(R0, rng.uniform(-1, 1, 3)), pose(rot(rng.normal(size=3), 2.4), rng.uniform(-1, 1, 3)), 1.3, 25, 3),
        "method1": (pose(R0, rng.uniform(-1, 1, 3)), pose(rot(rng.normal(size=3), 0.9), rng.uniform(-1, 1, 3)), 2.0, 12, 1),
        "same_R": (pose(R0, rng.uniform(-1, 1, 3)), pose(R0, rng.uniform(-1, 1, 3)), 0.8, 9, 5),
        "tiny": (pose(R0, rng.uniform(-1, 1, 3)), pose(R0 @ rot(ax, 3e-4), rng.uniform(-1, 1, 3)), 1.0, 9, 5),
        "near_pi": (pose(R0, rng.uniform(-1, 1, 3)), pose(R0 @ rot(ax, np.pi - 0.05), rng.uniform(-1, 1, 3)), 1.0, 21, 5),
        "pi_band": (pose(R0, rng.uniform(-1, 1, 3)), pose(R0 @ rot(ax, np.pi - 2e-3), rng.uniform(-1, 1, 3)), 1.0, 21, 3),
        "pi_exact": (pose(np.eye(3), np.zeros(3)), pose(np.diag([1.0, -1.0, -1.0]), np.ones(3)), 1.0, 11, 5),
    }
    out = {}
    for name, (Xs, Xe, Tf, N, method) in cases.items():
        r = planner.cartesian_trajectory(Xs, Xe, Tf, N, method)
        out.update({f"{name}_Xstart": Xs, f"{name}_Xend": Xe, f"{name}_args": np.array([Tf, N, method], np.float64)})
        for k in ("positions", "velocities", "accelerations", "orientations"):
            a = np.asarray(r[k])
            assert a.dtype == np.float32, (name, k, a.dtype)
            out[f"{name}_{k}"] = a
    np.savez(GOLD_DIR / "cartesian_trajectory.npz", **out)
    print("cartesian trajectory golden written")


def ik_golden() -> None:
    """iterative_inverse_kinematics of the unmodified reference (default mode): reachable
    targets T = FK(theta*) from seeds at increasing distance, an unreachable target that
    exhausts its iteration budget, and non-default damping / weights / step cap."""
    out = {}
    for robot, count in (("ur5", 10), ("iiwa14", 8)):
        proc, sm, dyn = load(robot)
        n = sm.S_list.shape[1]
        lims = limits_array(proc, n)
        rng = np.random.default_rng(12)
        tgt = rng.uniform(0.6 * lims[:, 0], 0.6 * lims[:, 1], (count, n))
        seeds = tgt + rng.uniform(-1, 1, (count, n)) * np.linspace(0.05, 0.9, count)[:, None]
        Td = np.stack([np.asarray(sm.forward_kinematics(t)) for t in tgt])
        Td[-1, :3, 3] += np.array([5.0, 0.0, 0.0])  # out of reach
        kw = [dict(max_iterations=300)] * count
        kw[1] = dict(max_iterations=300, damping=5e-2, step_cap=0.15)
        kw[2] = dict(max_iterations=300, weight_orientation=0.5, weight_position=2.0)
        kw[-1] = dict(max_iterations=60)
        th, ok, it = [], [], []
        for i in range(count):
            np.random.seed(100 + i)
            r = sm.iterative_inverse_kinematics(Td[i], seeds[i], **kw[i])
            th.append(np.asarray(r[0], np.float64))
            ok.append(bool(r[1]))
            it.append(int(r[2]))
        par = np.array([[k.get("max_iterations"), k.get("damping", 2e-2), k.get("step_cap", 0.3),
                         k.get("weight_orientation", 1.0), k.get("weight_position", 1.0)] for k in kw])
        # (self-contained: the reference's URDF loader picks the end-effector link from a set, so
        # M depends on PYTHONHASHSEED -- for the UR5 either the tool frame or the base link)
        out.update({f"{robot}_M": np.asarray(sm.M_list, np.float64), f"{robot}_S": np.asarray(sm.S_list, np.float64),
                    f"{robot}_T": Td, f"{robot}_seed": seeds, f"{robot}_params": par, f"{robot}_theta": np.stack(th),
                    f"{robot}_success": np.array(ok), f"{robot}_iterations": np.array(it),
                    f"{robot}_limits": lims})
    np.savez(GOLD_DIR / "inverse_kinematics.npz", **out)
    print("inverse kinematics golden written", {k: out[k] for k in out if k.endswith("iterations") or k.endswith("success")})


def ik_modes_golden() -> None:
    """iterative_inverse_kinematics of the unmodified reference with ``adaptive_tuning`` and / or
    ``backtracking`` (the modes smart_ / robust_inverse_kinematics switch on), same targets as
    ik_golden: (theta, success, iterations) per mode."""
    out = {}
    modes = {"adaptive": dict(adaptive_tuning=True), "backtracking": dict(backtracking=True),
             "both": dict(adaptive_tuning=True, backtracking=True)}
    for robot, count in (("ur5", 8), ("iiwa14", 8)):
        proc, sm, dyn = load(robot)
        n = sm.S_list.shape[1]
        lims = limits_array(proc, n)
        rng = np.random.default_rng(21)
        tgt = rng.uniform(0.6 * lims[:, 0], 0.6 * lims[:, 1], (count, n))
        seeds = tgt + rng.uniform(-1, 1, (count, n)) * np.linspace(0.05, 0.8, count)[:, None]
        Td = np.stack([np.asarray(sm.forward_kinematics(t)) for t in tgt])
        Td[-1, :3, 3] += np.array([5.0, 0.0, 0.0])  # out of reach
        budget = [300] * count
        budget[-1] = 60
        out.update({f"{robot}_M": np.asarray(sm.M_list, np.float64), f"{robot}_S": np.asarray(sm.S_list, np.float64),
                    f"{robot}_T": Td, f"{robot}_seed": seeds, f"{robot}_limits": lims,
                    f"{robot}_max_iterations": np.array(budget)})
        for name, kw in modes.items():
            th, ok, it = [], [], []
            for i in range(count):
                np.random.seed(200 + i)
                r = sm.iterative_inverse_kinematics(Td[i], seeds[i], max_iterations=budget[i], **kw)
                th.append(np.asarray(r[0], np.float64))
                ok.append(bool(r[1]))
                it.append(int(r[2]))
            out.update({f"{robot}_{name}_theta": np.stack(th), f"{robot}_{name}_success": np.array(ok),
                        f"{robot}_{name}_iterations": np.array(it)})
            print(robot, name, ok, it)
    np.savez(GOLD_DIR / "inverse_kinematics_modes.npz", **out)


def ik_front_golden() -> None:
    """smart_inverse_kinematics / robust_inverse_kinematics of the unmodified reference
    (kinematics/ik.py:327-598): targets reached from the first guess, targets that need the
    fall-back starts (NumPy's global generator seeded per call), one out of reach."""
    out = {}
    for robot, count in (("ur5", 8), ("iiwa14", 8)):
        proc, sm, dyn = load(robot)
        n = sm.S_list.shape[1]
        lims = limits_array(proc, n)
        sm.joint_limits = [tuple(r) for r in lims]
        rng = np.random.default_rng(33)
        tgt = rng.uniform(0.8 * lims[:, 0], 0.8 * lims[:, 1], (count, n))
        Td = np.stack([np.asarray(sm.forward_kinematics(t)) for t in tgt])
        Td[-1, :3, 3] += np.array([5.0, 0.0, 0.0])  # out of reach
        res = {k: [] for k in ("smart_theta", "smart_success", "smart_iterations", "smart_restarts", "robust_theta",
                               "robust_success", "robust_iterations", "robust_strategy", "robust_restarts")}
        # stagnation restarts of each run, counted by wrapping NumPy's randn (the solver's only use
        # of it): runs without any are the ones a solver with another noise source reproduces
        calls = {"n": 0}
        randn = np.random.randn

        def counting_randn(*a):
            calls["n"] += 1
            return randn(*a)

        np.random.randn = counting_randn
        for i in range(count):
            np.random.seed(300 + i)
            calls["n"] = 0
            th, ok, it = sm.smart_inverse_kinematics(Td[i], max_iterations=120)
            res["smart_theta"].append(np.asarray(th, np.float64))
            res["smart_success"].append(bool(ok))
            res["smart_iterations"].append(int(it))
            res["smart_restarts"].append(calls["n"])
            np.random.seed(400 + i)
            calls["n"] = 0
            th, ok, it, name = sm.robust_inverse_kinematics(Td[i], max_attempts=4, max_iterations=120)
            res["robust_theta"].append(np.asarray(th, np.float64))
            res["robust_success"].append(bool(ok))
            res["robust_iterations"].append(int(it))
            res["robust_strategy"].append(str(name))
            res["robust_restarts"].append(calls["n"])
        np.random.randn = randn
        out.update({f"{robot}_M": np.asarray(sm.M_list, np.float64), f"{robot}_S": np.asarray(sm.S_list, np.float64),
                    f"{robot}_T": Td, f"{robot}_limits": lims})
        for k, v in res.items():
            out[f"{robot}_{k}"] = np.array(v)
        print(robot, "smart", res["smart_success"], res["smart_iterations"], res["smart_restarts"])
        print(robot, "robust", res["robust_success"], res["robust_iterations"], res["robust_strategy"], res["robust_restarts"])
    np.savez(GOLD_DIR / "inverse_kinematics_front_ends.npz", **out)


API_CONTRACT_KEYS = (
    "ManipulatorDynamics.forward_dynamics", "ManipulatorDynamics.gravity_forces", "ManipulatorDynamics.inverse_dynamics",
    "ManipulatorDynamics.mass_matrix", "ManipulatorDynamics.velocity_quadratic_forces",
    "OptimizedTrajectoryPlanning.cartesian_trajectory", "OptimizedTrajectoryPlanning.forward_dynamics_trajectory",
    "OptimizedTrajectoryPlanning.inverse_dynamics_trajectory", "OptimizedTrajectoryPlanning.joint_trajectory",
    "SerialManipulator.forward_kinematics", "SerialManipulator.iterative_inverse_kinematics", "SerialManipulator.jacobian",
    "SerialManipulator.smart_inverse_kinematics", "SerialManipulator.robust_inverse_kinematics",
    "SerialManipulator.end_effector_velocity",
)


def api_contract_golden() -> None:
    """Hot-path subset of the reference's tests/data/api_contract_golden.json: per mirrored method
    the parameter names / optionality (from the reference classes themselves) and the pinned return
    type, dtype and shape."""
    import inspect
    import json

    import ManipulaPy.dynamics as dynamics_mod
    import ManipulaPy.kinematics as kinematics_mod

    classes = {"ManipulatorDynamics": dynamics_mod.ManipulatorDynamics,
               "SerialManipulator": kinematics_mod.SerialManipulator,
               "OptimizedTrajectoryPlanning": OptimizedTrajectoryPlanning}
    ref = json.loads((REF / "tests" / "data" / "api_contract_golden.json").read_text())
    out = {"_source": "reference tests/data/api_contract_golden.json (hot-path subset, oracle/gen_golden.py)"}
    for key in API_CONTRACT_KEYS:
        cls, meth = key.split(".")
        sig = inspect.signature(getattr(classes[cls], meth))
        out[key] = {
            "parameters": [{"name": q.name, "has_default": q.default is not inspect.Parameter.empty}
                           for q in sig.parameters.values() if q.name != "self"],
            "return": ref[key]["return"],
        }
    (GOLD_DIR / "api_contract.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("api contract subset:", len(out) - 1, "methods")


def singularity_golden() -> None:
    """Singularity.condition_number / singularity_analysis / near_singularity_detection of the
    unmodified reference (singularity/singularity_analysis.py:52-74, 246-305) at random and at
    singular configurations (stretched-out arm)."""
    import importlib

    sing_mod = importlib.import_module("ManipulaPy.singularity")
    out = {}
    for robot in ("ur5", "iiwa14"):
        proc, sm, dyn = load(robot)
        n = sm.S_list.shape[1]
        lims = limits_array(proc, n)
        rng = np.random.default_rng(44)
        th = rng.uniform(0.7 * lims[:, 0], 0.7 * lims[:, 1], (10, n))
        th[0] = 0.0  # home pose: singular for both arms
        th[1, 2:] = 0.0
        sg = sing_mod.Singularity(sm)
        out.update({f"{robot}_M": np.asarray(sm.M_list, np.float64), f"{robot}_S": np.asarray(sm.S_list, np.float64),
                    f"{robot}_thetas": th,
                    f"{robot}_condition_number": np.array([float(sg.condition_number(t)) for t in th]),
                    f"{robot}_singular": np.array([bool(sg.singularity_analysis(t)) for t in th]),
                    f"{robot}_near": np.array([bool(sg.near_singularity_detection(t)) for t in th])})
        print(robot, out[f"{robot}_condition_number"], out[f"{robot}_singular"])
    np.savez(GOLD_DIR / "singularity.npz", **out)


def body_kinematics_golden() -> None:
    """forward_kinematics / jacobian with frame="body" of the unmodified reference: the UR5 as
    loaded from its URDF, and a chain whose B_list is NOT Ad(M^-1) S_list (the reference takes
    any B_list; kinematics/serial_manipulator.py:75-95)."""
    from ManipulaPy.kinematics import SerialManipulator

    out = {}
    rng = np.random.default_rng(6)
    proc, sm, dyn = load("ur5")
    th = rng.uniform(-np.pi, np.pi, (9, 6))
    out.update(ur5_M=np.asarray(sm.M_list), ur5_S=np.asarray(sm.S_list), ur5_B=np.asarray(sm.B_list), ur5_theta=th,
               ur5_T=np.stack([np.asarray(sm.forward_kinematics(t, frame="body")) for t in th]),
               ur5_J=np.stack([np.asarray(sm.jacobian(t, frame="body")) for t in th]))
    # free-standing chain with one prismatic joint and independent body screws
    n = 5
    S, B = np.zeros((6, n)), np.zeros((6, n))
    for A in (S, B):
        for i in range(n):
            w = rng.normal(size=3)
            w /= np.linalg.norm(w)
            q = rng.uniform(-0.4, 0.4, 3)
            if i == 2:
                A[3:, i] = w
            else:
                A[:3, i] = w
                A[3:, i] = -np.cross(w, q)
    Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(Q) < 0:
        Q[:, 0] *= -1
    M = np.eye(4)
    M[:3, :3] = Q
    M[:3, 3] = rng.uniform(-0.5, 0.5, 3)
    sm2 = SerialManipulator(M_list=M, omega_list=S[:3], S_list=S, B_list=B)
    th2 = rng.uniform(-2, 2, (7, n))
    out.update(free_M=M, free_S=S, free_B=B, free_theta=th2,
               free_T=np.stack([np.asarray(sm2.forward_kinematics(t, frame="body")) for t in th2]),
               free_J=np.stack([np.asarray(sm2.jacobian(t, frame="body")) for t in th2]))
    np.savez(GOLD_DIR / "body_kinematics.npz", **out)
    print("body kinematics golden written")


def registry_trajectory_golden() -> None:
    """The registry seam (cuda_kernels/registry.py:828-867): outputs of the reference's own
    launcher with CUDA routing off, i.e. trajectory_cpu_fallback."""
    from ManipulaPy.cuda_kernels import registry

    rng = np.random.default_rng(2)
    out = {}
    cases = {"linear": (2.0, 64, 1), "method7": (1.5, 17, 7), "cubic": (0.9, 33, 3), "quintic": (2.0, 100, 5),
             "n1": (2.0, 1, 5), "tf0": (0.0, 9, 3), "tfneg": (-1.0, 4, 5)}
    for name, (Tf, N, method) in cases.items():
        s = rng.uniform(-2, 2, 6).astype(np.float32)
        e = rng.uniform(-2, 2, 6).astype(np.float32)
        pos, vel, acc = registry.execute_registered_kernel("trajectory.standard", s, e, Tf, N, method)
        out.update({f"{name}_start": s, f"{name}_end": e, f"{name}_args": np.array([Tf, N, method], np.float64),
                    f"{name}_positions": np.asarray(pos), f"{name}_velocities": np.asarray(vel),
                    f"{name}_accelerations": np.asarray(acc)})
    np.savez(GOLD_DIR / "registry_trajectory.npz", **out)
    print("registry trajectory golden written")


def id_trajectory_golden() -> None:
    out = {}
    for robot, npts in (("ur5", 24), ("iiwa14", 10)):
        proc, sm, dyn = load(robot)
        n = dyn.S_list.shape[1]
        lims = limits_array(proc, n)
        tl = np.tile(np.array([[-40.0, 35.0]]), (n, 1))  # finite: exercises the torque clip
        planner = OptimizedTrajectoryPlanning(sm, get_robot_urdf(robot), dyn, lims, tl, use_cuda=False)
        rng = np.random.default_rng(1)
        s, e = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        tr = planner.joint_trajectory(s, e, 2.0, 1000, 5)
        sel = np.linspace(0, 999, npts).astype(int)
        # float64 inputs (float32 arrays would make the reference run sin/cos in float32)
        th = np.asarray(tr["positions"], np.float64)[sel]
        dth = np.asarray(tr["velocities"], np.float64)[sel] * 3.0
        ddth = np.asarray(tr["accelerations"], np.float64)[sel] * 3.0
        ftip = np.array([0.5, -1.0, 0.25, 2.0, -1.5, 1.0])
        clear_caches(dyn)
        tau0 = np.asarray(planner.inverse_dynamics_trajectory(th, dth, ddth))
        clear_caches(dyn)
        tau1 = np.asarray(planner.inverse_dynamics_trajectory(th, dth, ddth, np.array([0.0, -9.81, 0.0]), ftip))
        assert tau0.dtype == np.float32
        out.update({f"{robot}_theta": th, f"{robot}_dtheta": dth, f"{robot}_ddtheta": ddth,
                    f"{robot}_torque_limits": tl, f"{robot}_tau_default": tau0,
                    f"{robot}_g1": np.array([0.0, -9.81, 0.0]), f"{robot}_ftip1": ftip,
                    f"{robot}_tau_g1_ftip1": tau1})
    np.savez(GOLD_DIR / "id_trajectory.npz", **out)
    print("id trajectory golden written")


def fd_trajectory_golden() -> None:
    out = {}
    for robot, N in (("iiwa14", 24), ("ur5", 16)):
        planner, dyn, lims = make_planner(robot)
        n = dyn.S_list.shape[1]
        rng = np.random.default_rng(4)
        th0 = rng.uniform(0.5 * lims[:, 0], 0.5 * lims[:, 1])
        dth0 = rng.uniform(-0.5, 0.5, n)
        tau = rng.uniform(-20, 20, (N, n))
        clear_caches(dyn)
        r = planner.forward_dynamics_trajectory(th0, dth0, tau, G_VEC, np.zeros((N, 6)), 1e-3, 1)
        out.update({f"{robot}_a_theta0": th0, f"{robot}_a_dtheta0": dth0, f"{robot}_a_tau": tau,
                    f"{robot}_a_args": np.array([1e-3, 1.0])})
        for k in ("positions", "velocities", "accelerations"):
            out[f"{robot}_a_{k}"] = np.asarray(r[k])
        # case b: intRes=2, non-zero wrench, start next to a joint limit with a hard push
        N2 = 8
        th0b = lims[:, 1] - 1e-4
        dth0b = np.full(n, 0.8)
        taub = rng.uniform(0, 40, (N2, n))
        ftb = rng.uniform(-2, 2, (N2, 6))
        clear_caches(dyn)
        r = planner.forward_dynamics_trajectory(th0b, dth0b, taub, np.array([0.0, -9.81, 0.0]), ftb, 5e-3, 2)
        out.update({f"{robot}_b_theta0": th0b, f"{robot}_b_dtheta0": dth0b, f"{robot}_b_tau": taub,
                    f"{robot}_b_ftip": ftb, f"{robot}_b_g": np.array([0.0, -9.81, 0.0]),
                    f"{robot}_b_args": np.array([5e-3, 2.0])})
        for k in ("positions", "velocities", "accelerations"):
            out[f"{robot}_b_{k}"] = np.asarray(r[k])
        out[f"{robot}_joint_limits"] = lims
    np.savez(GOLD_DIR / "fd_trajectory.npz", **out)
    print("fd trajectory golden written")


def main() -> None:
    ROBOT_DIR.mkdir(parents=True, exist_ok=True)
    GOLD_DIR.mkdir(parents=True, exist_ok=True)
    for robot in ("ur5", "iiwa14", "panda", "xarm6"):
        export_robot(robot)
    for robot in ("ur5", "panda", "iiwa14"):
        dynamics_golden(robot)
    trajectory_golden()
    body_kinematics_golden()
    cartesian_golden()
    ik_golden()
    ik_modes_golden()
    ik_front_golden()
    api_contract_golden()
    singularity_golden()
    registry_trajectory_golden()
    id_trajectory_golden()
    fd_trajectory_golden()


if __name__ == "__main__":
    main()
