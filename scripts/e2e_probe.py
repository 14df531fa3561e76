#!/usr/bin/env python3
"""End-to-end time of planner.trajectory_inverse_dynamics (NumPy in, NumPy float32 torques out) for
different numbers of pipelined device->host chunks, next to the bare PCIe copy of the same bytes."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from manipulapy_b200 import load_robot  # noqa: E402

rb = load_robot("ur5")
pl = rb.planner()
rng = np.random.default_rng(3)
B, N = 4096, 2441
s, e = rng.uniform(-np.pi, np.pi, (B, 6)), rng.uniform(-np.pi, np.pi, (B, 6))
res = {}
for chunks in (1, 4, 8, 16, 32, 64):
    pl.host_chunks = chunks
    out = None
    for _ in range(3):
        out = pl.trajectory_inverse_dynamics(s, e, 2.0, N, 5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        out = pl.trajectory_inverse_dynamics(s, e, 2.0, N, 5)
    res[f"chunks_{chunks}_ms"] = (time.perf_counter() - t0) / 10 * 1e3
pin = torch.empty(B * N * 6, dtype=torch.float32, pin_memory=True)
src = torch.empty(B * N * 6, dtype=torch.float32, device="cuda")
pin.copy_(src, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    pin.copy_(src, non_blocking=True)
torch.cuda.synchronize()
res["bare_copy_ms"] = (time.perf_counter() - t0) / 10 * 1e3
print(json.dumps(res))
