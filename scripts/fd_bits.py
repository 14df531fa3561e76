#!/usr/bin/env python3
"""Where do the float32 outputs of the rollout kernels differ over LONG rollouts?  The same 2,048 iiwa14 rollouts x N steps
through the three-warp kernel (batch of 2,048), the pair kernel (inside a batch of 16,384) and the single-warp kernel (inside
a batch of 24,576): number of differing entries per array, first step, ulps (tests: test_rollout_kernels_long_horizon).

    python scripts/fd_bits.py [N]
"""
import json, sys
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
    outs = {}
    for name, reps in (("small_2048", 1), ("mid_16384", 8), ("single_24576", 12)):
        out = ops.forward_dynamics_trajectory(h, th0.repeat(reps, 1), dth0.repeat(reps, 1), taum.repeat(reps, 1, 1), g, None,
                                              1e-3, 1, jl)
        outs[name] = [o[:B0].clone() for o in out]
        del out
    res = {}
    names = list(outs)
    for i in range(3):
        for j in range(i + 1, 3):
            a, b = outs[names[i]], outs[names[j]]
            key = names[i] + "_vs_" + names[j]
            res[key] = {}
            for lbl, x, y in zip(("pos", "vel", "acc"), a, b):
                ne = x.view(torch.int32) != y.view(torch.int32)
                cnt = int(ne.sum())
                first = None
                if cnt:
                    idx = ne.nonzero()
                    first = {"first_step": int(idx[:, 1].min()), "rollouts": int(idx[:, 0].unique().numel()),
                             "max_ulps": int((x.view(torch.int32)[ne] - y.view(torch.int32)[ne]).abs().max())}
                res[key][lbl] = {"differing": cnt, "of": x.numel(), **(first or {})}
    print(json.dumps(res, indent=1))

if __name__ == "__main__":
    main()
