#!/usr/bin/env python3
"""Per-kernel device timings at the BASELINE.json config sizes (CUDA events, L2 flushed
between iterations, inputs resident).  Prints one JSON object; used to fill DESIGN.md's
kernel table and profiles/.  Not the bench contract (that is bench.py)."""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

from manipulapy_b200 import _native, load_robot  # noqa: E402
from bench import ClockSampler  # noqa: E402  (nvidia-smi clocks / throttle reasons during the timed regions)

HBM = json.loads((REPO / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (REPO / "MEASURED_PEAKS.json").exists() else 6650.0


QUICK = "--quick" in sys.argv  # one launch per kernel (for ncu captures)


def timeit(fn, iters=10, warmup=3):
    flush = timeit.flush
    if QUICK:
        iters, warmup = 1, 0
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
        del out
    return float(np.median(ts)), float(min(ts))


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    timeit.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    main.sampler = ClockSampler(0).start()
    ops = _native.ops()
    res = {"hbm_peak_gbs": HBM, "device": torch.cuda.get_device_name(0)}
    g = [0.0, 0.0, -9.81]
    gen = torch.Generator(device=dev).manual_seed(0)

    def rand(*shape, lo=-1.0, hi=1.0, dtype=torch.float64):
        return lo + (hi - lo) * torch.rand(*shape, dtype=dtype, device=dev, generator=gen)

    def rec(name, units, sec, bytes_per_unit, flop_per_unit, unit):
        med, best = sec
        res[name] = {"units": units, "unit": unit, "ms": med * 1e3, "ms_best": best * 1e3,
                     "rate_per_s": units / med, "gbs": units * bytes_per_unit / med / 1e9,
                     "hbm_frac": units * bytes_per_unit / med / 1e9 / HBM,
                     "tflops": units * flop_per_unit / med / 1e12, "bytes_per_unit": bytes_per_unit,
                     "flop_per_unit": flop_per_unit}

    # fp64 / fp32 FMA peaks
    sink = torch.zeros(1, dtype=torch.float64, device=dev)
    for name, dt in (("fp64", 0), ("fp32", 1)):
        blocks, threads, iters = 148 * 8, 256, 1 << 15
        sec = timeit(lambda: ops.fma_peak(sink, dt, blocks, threads, iters), 5, 2)
        res[f"fma_peak_{name}_tflops"] = blocks * threads * iters * 16 / sec[1] / 1e12
    # fp64 issue rate under the operand patterns of the real kernels (csrc/robot.cu fma_peak modes):
    # instructions per clock per SM, against the 32 (= 64 lanes / 2) of the FMA pipe
    for mode, label in ((0, "shared_operands"), (1, "three_distinct_registers"), (2, "constant_bank_operand"),
                        (3, "dmul_dfma_pairs"), (4, "half_uniform_half_three_registers"),
                        (5, "three_registers_shared_middle_operand"), (6, "three_registers_two_shared_operands"),
                        (7, "three_registers_plus_one_imad_each"), (8, "constant_bank_operand_plus_one_imad_each")):
        blocks, threads, iters = 148 * 8, 256, 1 << 14
        sec = timeit(lambda: ops.fma_peak(sink, mode << 8, blocks, threads, iters), 5, 2)
        res[f"fp64_pattern_{label}_ginstr_per_s"] = blocks * threads * iters * 8 / sec[1] / 1e9

    # write-only HBM ceiling: hand-written 16-byte vector stores over 256 MiB .. 4 GiB (the round-1 figure
    # came from torch's memset on a buffer that half fits the 126 MB L2)
    stores = {}
    for gib in (0.25, 1, 2, 4):
        buf = torch.empty(int(gib * (1 << 30)), dtype=torch.uint8, device=dev)
        for mode, label in ((0, "st"), (1, "st.cs"), (2, "st.cg"), (3, "st.wt")):
            for blocks in (148 * 8, 148 * 32):
                sec = timeit(lambda: ops.store_peak(buf, mode, blocks), 5, 2)
                stores[f"{label}_{gib}GiB_{blocks}blocks"] = buf.numel() / sec[0] / 1e9
        sec = timeit(lambda: buf.zero_(), 5, 2)
        stores[f"torch_memset_{gib}GiB"] = buf.numel() / sec[0] / 1e9
        del buf
    res["store_gbs"] = stores
    res["store_peak_gbs"] = max(stores.values())

    if "--peaks-only" in sys.argv:
        res["clocks"] = main.sampler.stop()
        print(json.dumps(res, indent=1))
        return

    ur5 = load_robot("ur5", device=dev)
    h6 = ur5.dynamics.robot.handle
    B, N = 4096, 2441
    P = B * N
    s, e = rand(B, 6, lo=-np.pi, hi=np.pi), rand(B, 6, lo=-np.pi, hi=np.pi)
    jl = ur5.planner()._jl
    rec("joint_trajectory_ur5", P, timeit(lambda: ops.joint_trajectory(s, e, False, 2.0, N, 5, jl)), 72, 60, "points")
    rec("traj_rnea_fused_ur5", P, timeit(lambda: ops.trajectory_inverse_dynamics(h6, s, e, False, 2.0, N, 5, jl, g, None, None, False)), 24, 2130, "points")
    rec("traj_rnea_fused_with_traj_ur5", P, timeit(lambda: ops.trajectory_inverse_dynamics(h6, s, e, False, 2.0, N, 5, jl, g, None, None, True)), 96, 2130, "points")
    rec("traj_rnea_fused_f32compute_ur5", P, timeit(lambda: ops.trajectory_inverse_dynamics(h6, s, e, False, 2.0, N, 5, jl, g, None, None, False, True)), 24, 2130, "points")
    pos, vel, acc = ops.joint_trajectory(s, e, False, 2.0, N, 5, jl)
    p2, v2, a2 = pos.view(-1, 6), vel.view(-1, 6), acc.view(-1, 6)
    rec("rnea_f32io_ur5", P, timeit(lambda: ops.inverse_dynamics(h6, p2, v2, a2, g, None, None, None, True)), 96, 2070, "points")
    rec("rnea_f32io_f32compute_ur5", P, timeit(lambda: ops.inverse_dynamics(h6, p2, v2, a2, g, None, None, None, True, True)), 96, 2070, "points")
    p64, v64, a64 = p2.double(), v2.double(), a2.double()
    rec("rnea_f64io_ur5", P, timeit(lambda: ops.inverse_dynamics(h6, p64, v64, a64, g, None, None, None, False)), 192, 2070, "points")
    ft = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]
    rec("rnea_f64io_ftip_ur5", P, timeit(lambda: ops.inverse_dynamics(h6, p64, v64, a64, g, ft, None, None, False)), 192, 2070 + 690, "points")
    rec("gravity_f64io_ur5", P, timeit(lambda: ops.inverse_dynamics(h6, p64, None, None, g, None, None, None, False)), 96, 2070, "points")
    del pos, vel, acc, p64, v64, a64

    for name in ("iiwa14", "panda"):
        rb = load_robot(name, device=dev)
        h = rb.dynamics.robot.handle
        n = rb.num_joints
        Pk = 1_000_000
        lo = torch.from_numpy(rb.joint_limits[:, 0]).to(dev)
        hi = torch.from_numpy(rb.joint_limits[:, 1]).to(dev)
        th = lo + (hi - lo) * torch.rand(Pk, n, dtype=torch.float64, device=dev, generator=gen)
        rec(f"fk_jacobian_{name}", Pk, timeit(lambda: ops.fk_jacobian(h, th, True, True)), 8 * n + 128 + 48 * n, 170 * n, "configs")
        rec(f"fk_jacobian_f32_{name}", Pk, timeit(lambda: ops.fk_jacobian(h, th, True, True, True)), 8 * n + 64 + 24 * n, 170 * n, "configs")
        rec(f"fk_only_{name}", Pk, timeit(lambda: ops.fk_jacobian(h, th, True, False)), 8 * n + 128, 110 * n, "configs")
        rec(f"mass_matrix_{name}", Pk, timeit(lambda: ops.mass_matrix(h, th)), 8 * n + 8 * n * n, 64 * n + 100 * n + 53 * n * (n + 1) // 2, "configs")
        if name == "iiwa14":
            # batched DLS inverse kinematics: targets = FK of random configurations, seeds 0.3 rad away
            Pi = 200_000
            tgt = 0.5 * th[:Pi]
            Td = ops.fk_jacobian(h, tgt, True, False)[0]
            seed = tgt + 0.6 * (torch.rand(Pi, n, dtype=torch.float64, device=dev, generator=gen) - 0.5)
            lim = torch.from_numpy(rb.joint_limits.astype(np.float64))
            sol = ops.inverse_kinematics_dls(h, Td, seed, 1e-6, 1e-6, 400, 2e-2, 0.3, 1.0, 1.0, lim, 0)
            res["ik_dls_iiwa14_success_rate"] = float(sol[1].float().mean())
            res["ik_dls_iiwa14_mean_iterations"] = float(sol[2].float().mean())
            rec("ik_dls_iiwa14", Pi, timeit(lambda: ops.inverse_kinematics_dls(h, Td, seed, 1e-6, 1e-6, 400, 2e-2, 0.3, 1.0, 1.0, lim, 0), 3, 1),
                128 + 8 * n + 8 * n + 5, 1400 * float(sol[2].float().mean()), "targets")
            rec("ik_dls_one_phase_iiwa14", Pi, timeit(lambda: ops.inverse_kinematics_dls(h, Td, seed, 1e-6, 1e-6, 400, 2e-2, 0.3, 1.0, 1.0, lim, 0, False), 3, 1),
                128 + 8 * n + 8 * n + 5, 1400 * float(sol[2].float().mean()), "targets")
            # the same targets with adaptive tuning + line search (flags 3), as smart_ / robust_inverse_kinematics run it
            solm = ops.inverse_kinematics_dls(h, Td, seed, 1e-6, 1e-6, 400, 2e-2, 0.3, 1.0, 1.0, lim, 0, True, 3)
            res["ik_dls_modes_iiwa14_success_rate"] = float(solm[1].float().mean())
            res["ik_dls_modes_iiwa14_mean_iterations"] = float(solm[2].float().mean())
            rec("ik_dls_modes_iiwa14", Pi, timeit(lambda: ops.inverse_kinematics_dls(h, Td, seed, 1e-6, 1e-6, 400, 2e-2, 0.3, 1.0, 1.0, lim, 0, True, 3), 3, 1),
                128 + 8 * n + 8 * n + 5, (1400 + 5 * 350) * float(solm[2].float().mean()), "targets")
        dth, tau = rand(Pk, n), rand(Pk, n, lo=-20, hi=20)
        rec(f"forward_dynamics_{name}", Pk, timeit(lambda: ops.forward_dynamics(h, th, dth, tau, g, None, None)), 32 * n, 5300, "points")

    # Cartesian straight-line trajectories: 4096 pose pairs x 2441 steps, 72 B of float32 out per step
    Xs, _ = ops.fk_jacobian(h6, rand(B, 6, lo=-np.pi, hi=np.pi), True, False)
    Xe, _ = ops.fk_jacobian(h6, rand(B, 6, lo=-np.pi, hi=np.pi), True, False)
    rec("cartesian_trajectory", P, timeit(lambda: ops.cartesian_trajectory(Xs, Xe, 2.0, N, 5)), 72, 150, "points")
    del Xs, Xe

    iiwa = load_robot("iiwa14", device=dev)
    h7 = iiwa.dynamics.robot.handle
    lo = torch.from_numpy(iiwa.joint_limits[:, 0]).to(dev)
    hi = torch.from_numpy(iiwa.joint_limits[:, 1]).to(dev)
    jl7 = iiwa.planner()._jl
    # write-only and copy ceilings of this GPU for the roofline of the store-bound kernels
    big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    sec = timeit(lambda: big.zero_(), 5, 2)
    res["memset_gbs"] = big.numel() / sec[0] / 1e9
    src = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    sec = timeit(lambda: big.copy_(src), 5, 2)
    res["copy_gbs"] = 2 * big.numel() / sec[0] / 1e9
    del big, src
    for Bf in ((8192,) if QUICK else (65536, 8192)):
        Nf = 1000 if Bf == 65536 else 1000
        th0 = 0.5 * (lo + (hi - lo) * torch.rand(Bf, 7, dtype=torch.float64, device=dev, generator=gen))
        dth0 = rand(Bf, 7, lo=-0.5, hi=0.5)
        amp = torch.tensor([4.0, 4.0, 2.0, 2.0, 0.4, 0.2, 0.08], dtype=torch.float64, device=dev)
        nominal = iiwa.dynamics.gravity_forces(th0)
        taum = (nominal[:, None, :] + rand(Bf, Nf, 7, lo=-0.5, hi=0.5) * amp).float()
        rec(f"fd_rollout_iiwa14_B{Bf}", Bf * (Nf - 1),
            timeit(lambda: ops.forward_dynamics_trajectory(h7, th0, dth0, taum, g, None, 1e-3, 1, jl7), 3, 1),
            28 + 84, 5300, "steps")
        del taum
    res["clocks"] = main.sampler.stop()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
