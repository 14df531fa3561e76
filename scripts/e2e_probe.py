#!/usr/bin/env python3
"""experiment: where the end-to-end time of planner.trajectory_inverse_dynamics goes"""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from manipulapy_b200 import _host, _native, load_robot
dev = torch.device("cuda", 0)
rb = load_robot("ur5", device=dev)
pl = rb.planner()
rng = np.random.default_rng(3)
B, N = 4096, 2441
s = rng.uniform(-np.pi, np.pi, (B, 6)); e = rng.uniform(-np.pi, np.pi, (B, 6))

def tm(fn, k=5, w=2):
    for _ in range(w): r = fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3

print("api call            %.2f ms" % tm(lambda: pl.trajectory_inverse_dynamics(s, e, 2.0, N, 5)))
out = torch.empty((B, N, 6), dtype=torch.float32, pin_memory=True)
d = torch.empty((B, N, 6), dtype=torch.float32, device=dev)
print("raw D2H 240MB       %.2f ms" % tm(lambda: out.copy_(d, non_blocking=True)))
print("pinned alloc 240MB  %.2f ms" % tm(lambda: torch.empty((B, N, 6), dtype=torch.float32, pin_memory=True)))
def alloc_touch():
    t = torch.empty((B, N, 6), dtype=torch.float32, pin_memory=True); return t
print("to_device endpoints %.2f ms" % tm(lambda: (_host.to_device(s, dev), _host.to_device(e, dev))))
ops, h = _native.ops(), rb.dynamics.robot.handle
sd, ed = _host.to_device(s, dev), _host.to_device(e, dev)
g = [0.0, 0.0, -9.81]
for chunks in (1, 2, 4, 8, 16, 32):
    def run():
        def launch(lo, hi):
            return ops.trajectory_inverse_dynamics(h, sd[lo:hi], ed[lo:hi], False, 2.0, N, 5, pl._jl, g, None, None, False)[0]
        return _host.chunked_to_host(launch, B, (N, 6), torch.float32, dev, chunks=chunks)
    print("chunked chunks=%-3d   %.2f ms" % (chunks, tm(run)))
import os
print("cpus", os.cpu_count())
