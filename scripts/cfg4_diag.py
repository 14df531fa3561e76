#!/usr/bin/env python3
"""cfg 4 (iiwa14 rollouts): how far the kernels drift from the oracle over 1000 Euler steps, with the
oracle solving by LU (the reference's np.linalg.solve) and by the kernels' LDL^T; and what the survey's
literal torque distribution U(-20, 20) N m does at full size."""
import json, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from manipulapy_b200 import load_robot
from oracle import Oracle

def rel(a, b):
    b = np.asarray(b, np.float64); sc = np.maximum(1.0, np.abs(b).reshape(b.shape[0], -1).max(1))
    return float((np.abs(np.asarray(a, np.float64) - b).reshape(b.shape[0], -1).max(1) / sc).max())

rb = load_robot("iiwa14"); n = 7
o = Oracle(rb.S_list, rb.M, rb.Glist, rb.Mlist_per_link)
planner = rb.planner()
lo = torch.from_numpy(rb.joint_limits[:, 0]).cuda(); hi = torch.from_numpy(rb.joint_limits[:, 1]).cuda()
res = {}
for name, B in (("gravity_compensated", 8192), ("literal_u20", 65536)):
    N = 1000
    gen = torch.Generator(device="cuda").manual_seed(4)
    th0 = 0.5 * (lo + (hi - lo) * torch.rand(B, n, dtype=torch.float64, device="cuda", generator=gen))
    dth0 = torch.rand(B, n, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    if name == "literal_u20":
        tau = ((torch.rand(B, N, n, dtype=torch.float32, device="cuda", generator=gen) - 0.5) * 40.0)
    else:
        amp = torch.tensor([4.0, 4.0, 2.0, 2.0, 0.4, 0.2, 0.08], dtype=torch.float64, device="cuda")
        tau = (rb.dynamics.gravity_forces(th0)[:, None, :] + (torch.rand(B, N, n, dtype=torch.float64, device="cuda", generator=gen) - 0.5) * amp).float()
    r = planner.forward_dynamics_trajectory(th0, dth0, tau, [0, 0, -9.81], None, 1e-3, 1)
    acc = r["accelerations"]
    fin = torch.isfinite(acc).all(dim=2)
    bad = torch.nonzero(~fin.all(dim=1)).flatten()
    first_bad = [int((~fin[b]).float().argmax()) for b in bad[:8].tolist()]
    out = {"rollouts": B, "non_finite_rollouts": int(bad.numel()), "first_non_finite_steps": first_bad}
    sel = [0, 1, B // 2 + 1, B - 1]
    for solver, tag in ((True, "lu"), (2, "ldlt")):
        ref = o.forward_dynamics_trajectory(th0[sel].cpu().numpy(), dth0[sel].cpu().numpy(), tau[sel].double().cpu().numpy(),
                                            [0, 0, -9.81], None, 1e-3, 1, rb.joint_limits, analytic=solver)
        for steps in (50, 200, 1000):
            with np.errstate(invalid="ignore"):
                m = np.isfinite(ref["positions"][:, :steps]).all()
            out[f"rel_{tag}_{steps}"] = max(rel(r[k][sel, :steps].cpu().numpy().reshape(-1, n), np.nan_to_num(ref[k][:, :steps].reshape(-1, n), posinf=0, neginf=0))
                                            for k in ("positions", "velocities")) if m else None
    res[name] = out
    del tau, r
print(json.dumps(res, indent=1))
