#!/usr/bin/env python3
"""experiment: forward-dynamics rollout kernel variants (MPK_VARIANT / MPK_FD_THREADS), iiwa14"""
import json, os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from manipulapy_b200 import _native, load_robot
dev = torch.device("cuda", 0)
ops = _native.ops()
iiwa = load_robot("iiwa14", device=dev)
h7 = iiwa.dynamics.robot.handle
gen = torch.Generator(device=dev).manual_seed(0)
lo = torch.from_numpy(iiwa.joint_limits[:, 0]).to(dev); hi = torch.from_numpy(iiwa.joint_limits[:, 1]).to(dev)
jl7 = iiwa.planner()._jl
g = [0.0, 0.0, -9.81]
for Bf in [int(x) for x in sys.argv[1:]] or [65536, 8192]:
    Nf = int(os.environ.get("FD_STEPS", "1000"))
    th0 = 0.5 * (lo + (hi - lo) * torch.rand(Bf, 7, dtype=torch.float64, device=dev, generator=gen))
    dth0 = -0.5 + torch.rand(Bf, 7, dtype=torch.float64, device=dev, generator=gen)
    nominal = iiwa.dynamics.gravity_forces(th0)
    amp = torch.tensor([4.0, 4.0, 2.0, 2.0, 0.4, 0.2, 0.08], dtype=torch.float64, device=dev)
    taum = (nominal[:, None, :] + (-0.5 + torch.rand(Bf, Nf, 7, dtype=torch.float64, device=dev, generator=gen)) * amp).float()
    ts = []
    for it in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = ops.forward_dynamics_trajectory(h7, th0, dth0, taum, g, None, 1e-3, 1, jl7); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b)); del out
    print(f"variant {os.environ.get('MPK_VARIANT','0')} threads {os.environ.get('MPK_FD_THREADS','128')} B {Bf}: {min(ts[1:]):.3f} ms  {Bf*(Nf-1)/min(ts[1:])*1e3:.3e} steps/s", flush=True)
    del taum
