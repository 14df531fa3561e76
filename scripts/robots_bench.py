"""Fused trajectory + inverse-dynamics kernel across the bundled robots (device time per launch)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from manipulapy_b200 import _native, load_robot
dev = torch.device("cuda", 0)
ops = _native.ops()
gen = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); o = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)); del o
    return float(np.median(ts))
B, N = 4096, 2441
for name in ("ur5", "xarm6", "iiwa14", "panda"):
    rb = load_robot(name, device=dev)
    h, n = rb.dynamics.robot.handle, rb.num_joints
    lo = torch.from_numpy(rb.joint_limits[:, 0]).to(dev); hi = torch.from_numpy(rb.joint_limits[:, 1]).to(dev)
    s = lo + (hi - lo) * torch.rand(B, n, dtype=torch.float64, device=dev, generator=gen)
    e = lo + (hi - lo) * torch.rand(B, n, dtype=torch.float64, device=dev, generator=gen)
    jl = rb.planner()._jl
    g = [0.0, 0.0, -9.81]
    t64 = timeit(lambda: ops.trajectory_inverse_dynamics(h, s, e, False, 2.0, N, 5, jl, g, None, None, False))
    t32 = timeit(lambda: ops.trajectory_inverse_dynamics(h, s, e, False, 2.0, N, 5, jl, g, None, None, False, True))
    print(f"{name} n={n}: fused f64 {t64:.4f} ms {B*N/t64*1e3:.3e} pts/s | f32 {t32:.4f} ms {B*N/t32*1e3:.3e} pts/s", flush=True)
