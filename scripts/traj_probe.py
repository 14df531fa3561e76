#!/usr/bin/env python3
"""Device time of the trajectory-row kernel at the BASELINE shape (UR5 4096 x 2441) and a few others."""
import json, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from manipulapy_b200 import _native, load_robot
ops = _native.ops()
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for robot, B, N in (("ur5", 4096, 2441), ("iiwa14", 4096, 2441), ("ur5", 1 << 20, 8), ("ur5", 64, 100000)):
    rb = load_robot(robot, device=dev); n = rb.num_joints
    gen = torch.Generator(device=dev).manual_seed(0)
    s = (torch.rand(B, n, dtype=torch.float64, device=dev, generator=gen) * 2 - 1) * 3
    e = (torch.rand(B, n, dtype=torch.float64, device=dev, generator=gen) * 2 - 1) * 3
    jl = rb.planner()._jl
    for f32 in (False, True):
        for _ in range(3): ops.joint_trajectory(s, e, f32, 2.0, N, 5, jl)
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = ops.joint_trajectory(s, e, f32, 2.0, N, 5, jl); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b)); del out
        ms = float(np.median(ts))
        res[f"{robot}_{B}x{N}_{'f32in' if f32 else 'f64in'}"] = {"ms": ms, "gbs": B * N * n * 12 / ms / 1e6}
print(json.dumps(res, indent=1))
