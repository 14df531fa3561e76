#!/usr/bin/env python3
"""Golden vectors for EVERY robot of the reference's bundled database (test infrastructure).

Runs the unmodified reference (``/root/reference``, build container only) on each URDF that
``ManipulaPy.ManipulaPy_data.ROBOT_DATABASE`` lists and that its own loader turns into a serial
chain of at most 8 joints, and writes ``tests/golden/robot_zoo.npz``: per robot the constant pack
(``S_list``, ``M``, ``Glist``, ``Mlist_per_link``, limits) and, at a few random configurations,
``forward_kinematics`` / ``jacobian`` (kinematics/fk.py:39-86, jacobian.py:39-93),
``mass_matrix`` (dynamics/mass_matrix.py:16-99), ``gravity_forces`` /
``velocity_quadratic_forces`` (dynamics/forces.py:26-133), ``inverse_dynamics`` /
``forward_dynamics`` (dynamics/id_fd.py:16-83).  The chains differ in axis arrangements
(parallel, intersecting, offset wrists), which is what the kernels' Denavit-Hartenberg link
frames have to get right.

    python oracle/gen_robot_zoo.py
"""

from __future__ import annotations

import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import gen_golden as gg  # noqa: E402  (bootstraps the reference import)

import numpy as np  # noqa: E402
from ManipulaPy.ManipulaPy_data import ROBOT_DATABASE  # noqa: E402

N_CFG = 3


def main() -> None:
    out, names = {}, []
    for robot in sorted(ROBOT_DATABASE):
        try:
            proc, sm, dyn = gg.load(robot)
        except Exception as exc:  # noqa: BLE001
            print(f"skip {robot}: {type(exc).__name__}: {exc}")
            continue
        n = int(np.asarray(dyn.S_list).shape[1])
        if n < 1 or n > 8 or getattr(dyn, "Mlist_per_link", None) is None:
            print(f"skip {robot}: n={n}")
            continue
        S = np.asarray(dyn.S_list, np.float64)
        wn = np.linalg.norm(S[:3], axis=0)
        if not np.all((np.abs(wn - 1) < 1e-9) | (wn < 1e-12)):
            print(f"skip {robot}: non-unit screw axes")
            continue
        rng = np.random.default_rng(abs(hash(robot)) % (1 << 31) if False else sum(map(ord, robot)))
        lims = gg.limits_array(proc, n)
        lims = np.clip(lims, -2 * np.pi, 2 * np.pi)
        th = rng.uniform(lims[:, 0], lims[:, 1], size=(N_CFG, n))
        dth = rng.uniform(-1, 1, size=(N_CFG, n))
        ddth = rng.uniform(-1, 1, size=(N_CFG, n))
        tau = rng.uniform(-10, 10, size=(N_CFG, n))
        ft = np.zeros((N_CFG, 6))
        ft[-1] = [1.0, -2.0, 0.5, 3.0, -1.5, 0.75]
        r = {k: [] for k in ("fk", "jac", "mass", "g_forces", "c", "id", "fd")}
        for i in range(N_CFG):
            gg.clear_caches(dyn)
            r["fk"].append(np.asarray(sm.forward_kinematics(th[i])))
            r["jac"].append(np.asarray(sm.jacobian(th[i])))
            r["mass"].append(np.asarray(dyn.mass_matrix(th[i])))
            r["g_forces"].append(np.asarray(dyn.gravity_forces(th[i], gg.G_VEC)))
            r["c"].append(np.asarray(dyn.velocity_quadratic_forces(th[i], dth[i])))
            r["id"].append(np.asarray(dyn.inverse_dynamics(th[i], dth[i], ddth[i], gg.G_VEC, ft[i])))
            r["fd"].append(np.asarray(dyn.forward_dynamics(th[i], dth[i], tau[i], gg.G_VEC, ft[i])))
        names.append(robot)
        out.update({
            f"{robot}/S_list": S, f"{robot}/M": np.asarray(dyn.M_list, np.float64),
            f"{robot}/Glist": np.asarray(dyn.Glist, np.float64),
            f"{robot}/Mlist_per_link": np.asarray(dyn.Mlist_per_link, np.float64),
            f"{robot}/joint_limits": gg.limits_array(proc, n),
            f"{robot}/thetas": th, f"{robot}/dthetas": dth, f"{robot}/ddthetas": ddth,
            f"{robot}/taus": tau, f"{robot}/ftips": ft,
        })
        for k, v in r.items():
            out[f"{robot}/{k}"] = np.array(v, dtype=np.float64)
        print(f"{robot}: n={n}")
    out["robots"] = np.array(names)
    out["g"] = gg.G_VEC.copy()
    np.savez_compressed(gg.GOLD_DIR / "robot_zoo.npz", **out)
    print(f"robot zoo: {len(names)} robots -> tests/golden/robot_zoo.npz")


if __name__ == "__main__":
    main()
