#!/usr/bin/env python3
"""Do the three rollout kernels give the same bits over LONG rollouts?  The same 2,048 iiwa14 rollouts x N steps
through the three-warp kernel (batch of 2,048), the pair kernel (MPK_FD_SPLIT=2 in a second process, or the same 2,048
inside a batch of 16,000) and the single-warp kernel (inside a batch of 24,000).

    python scripts/fd_bits.py [N]
"""
import json
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from manipulapy_b200 import _native, load_robot  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    dev = torch.device("cuda", 0)
    ops = _native.ops()
    rb = load_robot("iiwa14", device=dev)
    n = rb.num_joints
    h, jl = rb.dynamics.robot.handle, rb.planner()._jl
    gen = torch.Generator(device=dev).manual_seed(4)
    lo = torch.from_numpy(rb.joint_limits[:, 0]).to(dev)
    hi = torch.from_numpy(rb.joint_limits[:, 1]).to(dev)
    B0 = 2048
    th0 = 0.5 * (lo + (hi - lo) * torch.rand(B0, n, dtype=torch.float64, device=dev, generator=gen))
    dth0 = torch.rand(B0, n, dtype=torch.float64, device=dev, generator=gen) - 0.5
    amp = torch.tensor([4.0, 4.0, 2.0, 2.0, 0.4, 0.2, 0.08], dtype=torch.float64, device=dev)
    taum = (rb.dynamics.gravity_forces(th0)[:, None, :]
            + (torch.rand(B0, N, n, dtype=torch.float64, device=dev, generator=gen) - 0.5) * amp).float()
    g = [0.0, 0.0, -9.81]
    res = {}
    outs = {}
    for name, reps in (("trio_2048", 1), ("pair_16384", 8), ("single_24576", 12)):
        out = ops.forward_dynamics_trajectory(h, th0.repeat(reps, 1), dth0.repeat(reps, 1), taum.repeat(reps, 1, 1), g, None,
                                              1e-3, 1, jl)
        outs[name] = [o[:B0].clone() for o in out]
        # every replica equals the first
        res[name + "_replicas_equal"] = all(bool((o.view(reps, B0, N, n).view(torch.int32) ==
                                                  o[:B0].view(torch.int32)).all()) for o in out)
        del out
    for name in ("pair_16384", "single_24576"):
        res["trio_equals_" + name] = [bool((a.view(torch.int32) == b.view(torch.int32)).all())
                                      for a, b in zip(outs["trio_2048"], outs[name])]
        res["max_abs_diff_vs_" + name] = [float((a.double() - b.double()).abs().nan_to_num(0).max())
                                          for a, b in zip(outs["trio_2048"], outs[name])]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
