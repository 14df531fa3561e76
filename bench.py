#!/usr/bin/env python3
"""bench.py -- RNEA-evaluated UR5 trajectory points/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], SURVEY.md 8d cfg 3): UR5, B = 4096 quintic joint
trajectories x N = 2441 steps = 9,998,336 points per GPU, Tf = 2, start/end ~ U(-pi, pi),
g = [0, 0, -9.81], Ftip = 0.  A step is one pass of the hot path over that batch:
trajectory generation (float64 time scaling -> float32 rows, clipped) and inverse dynamics
(float64 Newton-Euler recursion -> float32 torques, clipped) of every point, fused in one
kernel (`--mode fused`, default) or as the two reference calls (`--mode two_kernel`).
Multi-GPU: weak scaling, every rank owns its own 4096-trajectory shard (contiguous index
range of the global batch), no data-path collective.

`value`  : points/s with the endpoints resident in HBM, CUDA-event time per step summed over
           K steps, max over ranks; L2 is flushed between steps outside the timed regions.
`e2e`    : the same through the public host API (planner.trajectory_inverse_dynamics with
           NumPy endpoints in, NumPy float32 torques out): H2D + kernel + D2H per step.
`roofline`: dominant kernel's algorithmic HBM bytes / its CUDA-event duration against the
           measured copy bandwidth (MEASURED_PEAKS.json); the kernel is bound by the fp64
           pipe, so `roofline.fp64` carries the binding fractions: executed flops against an
           FMA peak measured in this same run, and executed fp64 instructions against the
           pipe's issue rate (peak, and with three distinct register operands per FMA).
`cpu_baseline` / `--impl reference`: the reference's algorithm (finite-difference Coriolis,
           sum_k Jk^T Gk Jk mass matrix; oracle/oracle.c literal port -- the reference itself
           is Python and cannot travel to the GPU box) on all host cores, on a bounded sample.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

ROBOT, B_TRAJ, N_STEPS, TF, METHOD = "ur5", 4096, 2441, 2.0, 5
FD_ROLLOUTS, FD_STEPS = 65536, 1000
METRIC, UNIT = "rnea_trajectory_points_per_s", "points/s"
# algorithmic work per point (DESIGN.md "Kernels")
FLOP_PER_POINT = 2070 + 60            # textbook fp64 Newton-Euler recursion (n = 6) + time scaling
# what the fused kernel actually executes per point (ncu, profiles/r1_ncu_traj_rnea_dh.md):
# 514 DFMA + 181 DMUL + 41 DADD = 736 fp64 instructions = 1250 flop (Denavit-Hartenberg frames,
# centre-of-mass wrench form)
FP64_INSTR_PER_POINT = {"fused": 736, "two_kernel": 712}
FLOP_EXECUTED_PER_POINT = {"fused": 1250, "two_kernel": 1226}
BYTES_FUSED = 6 * 4                   # float32 torque row out; endpoints amortised over 2441 points
BYTES_RNEA = 3 * 6 * 4 + 6 * 4        # float32 theta, dtheta, ddtheta in; float32 torque out
BYTES_TRAJ = 3 * 6 * 4                # float32 pos, vel, acc out


_RESULT: list = []  # the JSON line, printed by main() once stdout is restored


def _env_int(name: str, default: int) -> int:
    return int(os.environ.get(name, default))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _measured_peaks() -> tuple[float, str]:
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic(kernel: str):
    p = REPO / "profiles" / "traffic.json"
    if p.exists():
        return json.loads(p.read_text()).get(kernel)
    return None


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores
# ------------------------------------------------------------------------------------------
def cpu_reference_rate(sample_traj: int, steps: int, warmup: int, seed: int = 3):
    """points/s of trajectory generation + literal (finite-difference) inverse dynamics."""
    from manipulapy_b200.robots import load_robot
    from oracle import Oracle, oracle_lib

    rb = load_robot(ROBOT)
    cores = os.cpu_count() or 1
    oracle_lib.set_threads(cores)
    o = Oracle(rb.S_list, rb.M, rb.Glist, rb.Mlist_per_link)
    rng = np.random.default_rng(seed)
    s = rng.uniform(-np.pi, np.pi, (sample_traj, 6))
    e = rng.uniform(-np.pi, np.pi, (sample_traj, 6))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        tr = Oracle.joint_trajectory(s, e, TF, N_STEPS, METHOD, rb.joint_limits)
        o.inverse_dynamics_trajectory(tr["positions"].reshape(-1, 6), tr["velocities"].reshape(-1, 6),
                                      tr["accelerations"].reshape(-1, 6), analytic=False)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    pts = sample_traj * N_STEPS
    return pts * len(times) / sum(times), cores, pts, sum(times) / len(times)


def run_reference(args) -> None:
    rank = _env_int("RANK", 0)
    if rank != 0:
        return
    sample = max(1, args.ref_traj)
    rate, cores, pts, sec = cpu_reference_rate(sample, args.steps, args.warmup)
    desc = (f"{sample} of {B_TRAJ} trajectories x {N_STEPS} steps per step ({pts} points); "
            "oracle/oracle.c literal port of the reference algorithm (the Python reference cannot travel)")
    _RESULT.append(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the same workload as the GPU arm's line (each step a bounded sample of it, see cpu_baseline.sample)
        "config": {"workload": f"UR5 {B_TRAJ} trajectories x {N_STEPS} steps per GPU, quintic joint_trajectory + "
                               "inverse_dynamics_trajectory (BASELINE.json configs[2])",
                   "robot": ROBOT, "trajectories_per_gpu": B_TRAJ, "steps_per_trajectory": N_STEPS,
                   "points_per_gpu": B_TRAJ * N_STEPS, "mode": "cpu reference algorithm, sampled",
                   "outputs": "float32 torques (B, N, 6)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from manipulapy_b200 import _native, load_robot, shard_range

    rank, world, local = _env_int("RANK", 0), _env_int("WORLD_SIZE", 1), _env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one process per GPU: stay on the CPU cores (and NUMA node) next to this GPU so the pinned
    # result buffers of the end-to-end path do not cross the socket link
    from manipulapy_b200 import bind_host_to_device
    numa_cpus = bind_host_to_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops = _native.ops()
    rb = load_robot(ROBOT, device=dev)
    planner = rb.planner()
    handle = rb.dynamics.robot.handle
    jl = planner._jl

    # this rank's contiguous shard of the global (world * 4096)-trajectory batch
    lo, hi = shard_range(world * B_TRAJ, world, rank)
    rng = np.random.default_rng(3)
    start_all = rng.uniform(-np.pi, np.pi, (world * B_TRAJ, 6))
    end_all = rng.uniform(-np.pi, np.pi, (world * B_TRAJ, 6))
    s_host, e_host = start_all[lo:hi].copy(), end_all[lo:hi].copy()
    s, e = torch.from_numpy(s_host).to(dev), torch.from_numpy(e_host).to(dev)
    B = hi - lo
    P = B * N_STEPS
    g = [0.0, 0.0, -9.81]

    if args.mode == "fused":
        def step():
            return ops.trajectory_inverse_dynamics(handle, s, e, False, TF, N_STEPS, METHOD, jl, g, None, None, False)[0]
        launches_per_step, dom_kernel, dom_bytes = 2, "traj_rnea_kernel<6,false>", BYTES_FUSED  # + time-scaling table kernel
    else:
        def step():
            pos, vel, acc = ops.joint_trajectory(s, e, False, TF, N_STEPS, METHOD, jl)
            return ops.inverse_dynamics(handle, pos.view(-1, 6), vel.view(-1, 6), acc.view(-1, 6), g, None, None, None, True)
        launches_per_step, dom_kernel, dom_bytes = 3, "rnea_kernel<6,false>", BYTES_RNEA  # table + trajectory + RNEA

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = fn()
            b.record()
            evs.append((a, b))
            del out
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs) / 1e3  # seconds over `steps`

    sampler = ClockSampler(local).start() if rank == 0 else None
    t_dev = timed(step, args.steps, args.warmup)

    # dominant kernel alone (two_kernel mode: the RNEA launch; fused: the step is that kernel)
    if args.mode == "fused":
        t_dom = t_dev
    else:
        pos, vel, acc = ops.joint_trajectory(s, e, False, TF, N_STEPS, METHOD, jl)
        t_dom = timed(lambda: ops.inverse_dynamics(handle, pos.view(-1, 6), vel.view(-1, 6), acc.view(-1, 6), g,
                                                   None, None, None, True), args.steps, 1)
        del pos, vel, acc

    # end to end through the public host API: NumPy endpoints in, NumPy float32 torques out
    def e2e_step():
        return planner.trajectory_inverse_dynamics(s_host, e_host, TF, N_STEPS, METHOD)

    # warm-up in the steady-state pattern of the timed loop (the caller holds the previous result
    # while the next one is produced, so two pinned result buffers are in rotation; the first
    # use of each is a ~100 ms cudaHostAlloc that torch's host allocator then caches)
    out = None
    for _ in range(max(3, args.warmup)):
        out = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = e2e_step()
    barrier()
    t_e2e = time.perf_counter() - t0
    assert out.shape == (B, N_STEPS, 6) and out.dtype == np.float32
    clocks = sampler.stop() if sampler else None

    # fp64 FMA peak measured in this run (register-resident dependent chains, 8 per thread), and
    # the issue rate of FMAs with three DISTINCT register operands (what rigid-body algebra is
    # made of): the register file feeds those at about 72 % of the pipe's peak rate
    sink = torch.zeros(1, dtype=torch.float64, device=dev)
    blocks, threads, iters = 148 * 8, 256, 1 << 15

    def fma_rate(mode):
        ops.fma_peak(sink, mode << 8, blocks, threads, 1024)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.fma_peak(sink, mode << 8, blocks, threads, iters)
            b.record()
            torch.cuda.synchronize()
            best = max(best, blocks * threads * iters * 8 / (a.elapsed_time(b) / 1e3))
        return best  # fp64 thread-instructions / s

    fp64_instr_peak = fma_rate(0)
    fp64_instr_3reg = fma_rate(1)
    fp64_peak_tf = 2 * fp64_instr_peak / 1e12

    # PCIe device -> host rate of this box (pinned memory), the ceiling of the e2e number
    pin = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
    src = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pin.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    pin.copy_(src, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    d2h_gbs = pin.numel() / (a.elapsed_time(b) / 1e3) / 1e9
    # ... and with every rank copying at the same time: the GPUs of one box share PCIe uplinks, so
    # the per-GPU rate drops (8 GPUs: 57 -> 12-18 GB/s, profiles/r1_d2h_n8.json), and the e2e time
    # is the slowest rank's
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4):
        pin.copy_(src, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    d2h_conc = torch.tensor([4 * pin.numel() / (a.elapsed_time(b) / 1e3) / 1e9], dtype=torch.float64, device=dev)
    d2h_conc_sum = d2h_conc.clone()
    if world > 1:
        dist.all_reduce(d2h_conc, op=dist.ReduceOp.MIN)
        dist.all_reduce(d2h_conc_sum, op=dist.ReduceOp.SUM)
    d2h_conc_min, d2h_conc_sum = float(d2h_conc), float(d2h_conc_sum)
    del pin, src

    # second metric of BASELINE.json: forward-dynamics rollout steps/s (configs[3]: iiwa14, 65,536
    # shooting trajectories x 1000 Euler steps per GPU, float32 torque rows resident in HBM)
    t_fd, fd_steps, t_fd_small = 0.0, 0, 0.0
    if not args.no_fd:
        iiwa = load_robot("iiwa14", device=dev)
        h7, jl7 = iiwa.dynamics.robot.handle, iiwa.planner()._jl
        Bf, Nf = FD_ROLLOUTS, FD_STEPS
        gen = torch.Generator(device=dev).manual_seed(4 + rank)
        lo7 = torch.from_numpy(iiwa.joint_limits[:, 0]).to(dev)
        hi7 = torch.from_numpy(iiwa.joint_limits[:, 1]).to(dev)
        th0 = 0.5 * (lo7 + (hi7 - lo7) * torch.rand(Bf, 7, dtype=torch.float64, device=dev, generator=gen))
        dth0 = torch.rand(Bf, 7, dtype=torch.float64, device=dev, generator=gen) - 0.5
        amp = torch.tensor([4.0, 4.0, 2.0, 2.0, 0.4, 0.2, 0.08], dtype=torch.float64, device=dev)
        taum = (iiwa.dynamics.gravity_forces(th0)[:, None, :]
                + (torch.rand(Bf, Nf, 7, dtype=torch.float64, device=dev, generator=gen) - 0.5) * amp).float()
        t_fd = timed(lambda: ops.forward_dynamics_trajectory(h7, th0, dth0, taum, g, None, 1e-3, 1, jl7), 3, 1)
        fd_steps = Bf * (Nf - 1) * 3
        # one GPU's share of the same 65,536 rollouts on 8 GPUs (strong scaling): each step split across a warp pair
        Bs = Bf // 8
        t_fd_small = timed(lambda: ops.forward_dynamics_trajectory(h7, th0[:Bs], dth0[:Bs], taum[:Bs], g, None, 1e-3, 1,
                                                                   jl7), 3, 1) / 3
        del taum

    red = torch.tensor([t_dev, t_dom, t_e2e, t_fd], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    t_dev, t_dom, t_e2e, t_fd = (float(x) for x in red.cpu())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_pts = world * P * args.steps
    value = total_pts / t_dev
    hbm_peak, peak_src = _measured_peaks()
    dom_s = t_dom / args.steps
    achieved_gbs = dom_bytes * P / dom_s / 1e9
    achieved_tf = FLOP_EXECUTED_PER_POINT[args.mode] * P / dom_s / 1e12
    instr_rate = FP64_INSTR_PER_POINT[args.mode] * P / dom_s
    cpu_rate, cores, cpu_pts, _ = cpu_reference_rate(args.cpu_sample_traj, 1, 0) if args.gpus == 1 and not args.no_cpu else (None, None, None, None)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"UR5 {B_TRAJ} trajectories x {N_STEPS} steps per GPU, quintic joint_trajectory + "
                               "inverse_dynamics_trajectory (BASELINE.json configs[2])",
                   "robot": ROBOT, "trajectories_per_gpu": B, "steps_per_trajectory": N_STEPS, "points_per_gpu": P,
                   "mode": args.mode, "outputs": "float32 torques (B, N, 6)",
                   "l2": "256 MiB buffer written between timed steps (outside the event pairs); "
                         "each step also writes 240 MB > 126 MB L2",
                   "sharding": f"contiguous trajectory ranges, {world} rank(s), no data-path collective"},
        "clocks": clocks,
        "e2e": {"value": world * P * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(2 * B * 6 * 8),
                "d2h_bytes_per_step": int(P * 6 * 4), "api": "OptimizedTrajectoryPlanning.trajectory_inverse_dynamics",
                "pcie_d2h_gbs_measured": d2h_gbs, "pcie_d2h_gbs_all_ranks_together_min": d2h_conc_min,
                "pcie_d2h_gbs_all_ranks_together_sum": d2h_conc_sum, "host_cpus_bound_to": numa_cpus,
                # ceiling of the e2e number: every rank's 240 MB result crosses PCIe at the
                # slowest rank's share of the box's uplinks
                "pcie_bound_points_per_s": world * d2h_conc_min * 1e9 / (6 * 4)},
        "gpu_launches": launches_per_step * args.steps,  # in the timed region of `value`
        "roofline": {"bound": "hbm", "kernel": dom_kernel, "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_peak, "traffic": _traffic(dom_kernel), "peak_source": peak_src,
                     "algorithmic_bytes_per_point": dom_bytes, "kernel_ms": dom_s * 1e3,
                     "binding": "fp64_pipe",
                     "fp64": {"achieved": achieved_tf, "peak": fp64_peak_tf, "unit": "TFLOP/s",
                              "frac": achieved_tf / fp64_peak_tf,
                              "flop_per_point": FLOP_EXECUTED_PER_POINT[args.mode],
                              "flop_per_point_textbook": FLOP_PER_POINT,
                              "textbook_equivalent_tflops": FLOP_PER_POINT * P / dom_s / 1e12,
                              "pipe_issue_frac": instr_rate / fp64_instr_peak,
                              "pipe_issue_frac_vs_3_register_operand_rate": instr_rate / fp64_instr_3reg,
                              "fp64_instr_per_point": FP64_INSTR_PER_POINT[args.mode],
                              "instr_peak_per_s": fp64_instr_peak, "instr_3reg_rate_per_s": fp64_instr_3reg,
                              "peak_source": "mpk_fma_peak measured in this run (mode 0: shared operands = "
                                             "pipe peak; mode 1: three distinct register operands)"}},
    }
    if t_fd > 0:
        line["fd_rollout"] = {
            "metric": "fd_rollout_steps_per_s", "value": world * fd_steps / t_fd, "unit": "steps/s",
            "ms_per_launch": t_fd / 3 * 1e3, "gpu_launches": 3,
            "ms_per_launch_8192_rollouts": t_fd_small * 1e3,  # rank 0's time (not reduced over ranks)
            "config": {"workload": f"iiwa14 {FD_ROLLOUTS} rollouts x {FD_STEPS} Euler steps per GPU, CRBA mass matrix + "
                                   "LDL^T solve per step (BASELINE.json configs[3])", "dt": 1e-3, "intRes": 1,
                       "taumat": "float32 (B, N, 7) resident in HBM", "outputs": "3 x float32 (B, N, 7)"}}
    if cpu_rate is not None:
        line["cpu_baseline"] = {
            "value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{args.cpu_sample_traj} of {B_TRAJ} trajectories x {N_STEPS} steps ({cpu_pts} points), "
                      "oracle/oracle.c literal port of the reference algorithm, all host threads"}
    _RESULT.append(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--mode", choices=["fused", "two_kernel"], default="fused")
    ap.add_argument("--ref-traj", type=int, default=16, help="trajectories per step of the reference arm")
    ap.add_argument("--cpu-sample-traj", type=int, default=256, help="trajectories of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fd", action="store_true", help="skip the forward-dynamics rollout metric")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly ONE JSON line: anything a library prints there meanwhile (e.g. NCCL's
    # version banner) is diverted to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        (run_reference if args.impl == "reference" else run_ours)(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if _RESULT:
        print(_RESULT[0], flush=True)


if __name__ == "__main__":
    main()
