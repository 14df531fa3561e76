#!/usr/bin/env python3
"""Forward-dynamics rollouts alone (iiwa14): device time per launch for a given batch / step
count.  Used for ncu source-level captures of ``fd_rollout_kernel`` and for variant sweeps.

    python scripts/fd_probe.py [B] [N] [iters]
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

from manipulapy_b200 import _native, load_robot  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    robot = sys.argv[4] if len(sys.argv) > 4 else "iiwa14"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ops = _native.ops()
    rb = load_robot(robot, device=dev)
    n = rb.num_joints
    h, jl = rb.dynamics.robot.handle, rb.planner()._jl
    gen = torch.Generator(device=dev).manual_seed(4)
    lo = torch.from_numpy(rb.joint_limits[:, 0]).to(dev)
    hi = torch.from_numpy(rb.joint_limits[:, 1]).to(dev)
    th0 = 0.5 * (lo + (hi - lo) * torch.rand(B, n, dtype=torch.float64, device=dev, generator=gen))
    dth0 = torch.rand(B, n, dtype=torch.float64, device=dev, generator=gen) - 0.5
    amp = torch.tensor([4.0, 4.0, 2.0, 2.0, 0.4, 0.2, 0.08, 0.08][:n], dtype=torch.float64, device=dev)
    taum = (rb.dynamics.gravity_forces(th0)[:, None, :]
            + (torch.rand(B, N, n, dtype=torch.float64, device=dev, generator=gen) - 0.5) * amp).float()
    g = [0.0, 0.0, -9.81]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for it in range(iters + 1):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = ops.forward_dynamics_trajectory(h, th0, dth0, taum, g, None, 1e-3, 1, jl)
        b.record()
        torch.cuda.synchronize()
        if it:
            ts.append(a.elapsed_time(b))
    chk = [float(torch.nan_to_num(o.double(), nan=0.0, posinf=0.0, neginf=0.0).sum()) for o in out]
    bad = [int((~torch.isfinite(o)).any(dim=2).any(dim=1).sum()) for o in out]  # rollouts with a non-finite row
    print(json.dumps({"robot": robot, "B": B, "N": N, "ms": ts, "ms_min": min(ts),
                      "steps_per_s": B * (N - 1) / (min(ts) * 1e-3), "checksum": chk, "rollouts_with_nonfinite_rows": bad}))


if __name__ == "__main__":
    main()
