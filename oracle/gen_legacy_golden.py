#!/usr/bin/env python3
"""Golden vectors of the reference's LEGACY dynamics path: a ``ManipulatorDynamics``
built by hand without ``Mlist_per_link`` (as tests/test_dynamics.py:59-67 does) falls back to
``_mass_matrix_legacy`` (dynamics/mass_matrix.py:101-132) and ``_gravity_forces_legacy``
(dynamics/forces.py:135-154); Coriolis forces, inverse and forward dynamics are then built on those.
The reference itself documents this path as incorrect physics; it is reproduced for callers that
still construct the object that way.

TEST INFRASTRUCTURE ONLY (runs the unmodified reference in the build container):

    python oracle/gen_legacy_golden.py      ->  tests/golden/legacy_dynamics.npz
"""
import importlib.util
import warnings
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
_spec = importlib.util.spec_from_file_location("gen_golden", REPO / "oracle" / "gen_golden.py")
gg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gg)

from ManipulaPy.dynamics import ManipulatorDynamics  # noqa: E402


def main():
    out = {}
    warnings.filterwarnings("ignore")
    for robot in ("ur5", "iiwa14"):
        with np.load(REPO / "manipulapy_b200" / "robots" / f"{robot}.npz") as d:
            S, M, G = d["S_list"], d["M"], d["Glist"]
        n = S.shape[1]
        rng = np.random.default_rng(41 + n)
        # a general (non-symmetric-looking but symmetric) inertia on one link exercises the 6x6 products
        G = G.copy()
        A = rng.uniform(-0.05, 0.05, (6, 6))
        G[2] = G[2] + A + A.T
        dyn = ManipulatorDynamics(M_list=M, omega_list=None, r_list=None, b_list=None, S_list=S, B_list=None, Glist=G)
        P = 6
        th, dth, ddth = rng.uniform(-2, 2, (P, n)), rng.uniform(-1.5, 1.5, (P, n)), rng.uniform(-3, 3, (P, n))
        tau, ft = rng.uniform(-10, 10, (P, n)), rng.uniform(-4, 4, (P, 6))
        g = np.array([0.3, -0.2, -9.81])
        res = {k: [] for k in ("mass", "grav", "cor", "id", "fd")}
        for p in range(P):
            dyn._mass_matrix_cache.clear()
            dyn._mass_matrix_derivative_cache.clear()
            res["mass"].append(np.asarray(dyn.mass_matrix(th[p])))
            res["grav"].append(np.asarray(dyn.gravity_forces(th[p], g)))
            res["cor"].append(np.asarray(dyn.velocity_quadratic_forces(th[p], dth[p])))
            res["id"].append(np.asarray(dyn.inverse_dynamics(th[p], dth[p], ddth[p], g, ft[p])))
            res["fd"].append(np.asarray(dyn.forward_dynamics(th[p], dth[p], tau[p], g, ft[p])))
        out.update({f"{robot}_{k}": np.stack(v) for k, v in res.items()})
        out.update({f"{robot}_S": S, f"{robot}_M": M, f"{robot}_G": G, f"{robot}_th": th, f"{robot}_dth": dth,
                    f"{robot}_ddth": ddth, f"{robot}_tau": tau, f"{robot}_ft": ft, f"{robot}_g": g})
        print(robot, "legacy golden:", P, "configurations; |M| max", float(np.abs(out[f"{robot}_mass"]).max()))
    np.savez(REPO / "tests" / "golden" / "legacy_dynamics.npz", **out)


if __name__ == "__main__":
    main()
