"""Device time of the Cartesian straight-line trajectory kernel (4096 pose pairs x 2441 steps):  python scripts/cart_probe.py"""
import sys, json, torch, numpy as np
sys.path.insert(0, '.')
from manipulapy_b200 import _native, load_robot
ops = _native.ops(); dev = torch.device('cuda:0')
ur5 = load_robot('ur5', device=dev); h6 = ur5.dynamics.robot.handle
g = torch.Generator(device=dev).manual_seed(0)
B, N = 4096, 2441
th = lambda: (torch.rand(B, 6, dtype=torch.float64, device=dev, generator=g) * 2 - 1) * np.pi
Xs, _ = ops.fk_jacobian(h6, th(), True, False); Xe, _ = ops.fk_jacobian(h6, th(), True, False)
ts = []
for i in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = ops.cartesian_trajectory(Xs, Xe, 2.0, N, 5); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b)); del out
print(json.dumps({"cartesian_ms": ts}))
