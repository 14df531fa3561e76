#!/usr/bin/env python3
"""bench.py -- RNEA-evaluated UR5 trajectory points/s (BASELINE.json metric), every BASELINE config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

HEADLINE (BASELINE.json configs[2], SURVEY.md 8d cfg 3): UR5, B = 4096 quintic joint
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
           NumPy endpoints in, NumPy float32 torques out): H2D + kernel + D2H per step.  At
           N > 1 the global batch is sharded in proportion to each rank's measured share of
           the box's PCIe uplinks (the GPUs of one box do not have equal device->host rates).
`roofline`: dominant kernel's algorithmic HBM bytes / its CUDA-event duration against the
           measured copy bandwidth (MEASURED_PEAKS.json); the kernel is bound by the fp64
           pipe, so `roofline.fp64` carries the binding fractions.
`configs`: one entry per remaining BASELINE.json config, each with its own device time,
           roofline fraction and clock record:
             cfg1  UR5 N = 1000 joint_trajectory + inverse_dynamics_trajectory (two calls)
             cfg2  FK + space Jacobian over 1 M configurations (iiwa14: true 7-DOF; Panda: the
                   reference's 8-DOF chain)                      + mass matrix, trajectory rows
             cfg4  iiwa14 forward-dynamics rollouts, 65,536 x 1000 Euler steps
             cfg5  fused trajectory + inverse dynamics over 1e6 .. 1e9 points, STRONG scaling:
                   the trajectories are split by `shard_range` over the N ranks and the float32
                   torque rows are gathered on rank 0 INSIDE the timed region -- by the kernel
                   itself storing into rank 0's peer-mapped buffer over NVLink (`PeerRows`), and,
                   as the baseline, by a chunked NCCL send / recv pipeline on a second stream.
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
# what the fused kernel executes per point (scripts/sass_census.py on the main path; ncu agrees):
# fp64 instructions (DFMA + DMUL + DADD) and the flops they stand for
# (ncu, profiles/r2_traj_rnea_instruction_mix.csv: 439 DFMA + 138 DMUL + 48 DADD per point with the UR
# family's link-geometry kernels; round 1's general kernel executed 514 + 181 + 41 = 736)
FP64_INSTR_PER_POINT = {"fused": 625, "two_kernel": 601}
FLOP_EXECUTED_PER_POINT = {"fused": 1064, "two_kernel": 1034}
BYTES_FUSED = 6 * 4                   # float32 torque row out; endpoints amortised over 2441 points
BYTES_RNEA = 3 * 6 * 4 + 6 * 4        # float32 theta, dtheta, ddtheta in; float32 torque out
BYTES_TRAJ = 3 * 6 * 4                # float32 pos, vel, acc out
# the unmodified Python reference, measured in the build container (one
# ManipulatorDynamics.inverse_dynamics call on the UR5 takes 61 ms on one core)
PY_REFERENCE_POINTS_PER_S_PER_CORE = 1.0 / 0.061

_RESULT: list = []  # the JSON line, printed by main() once stdout is restored


def _env_int(name: str, default: int) -> int:
    return int(os.environ.get(name, default))


def headline_config(world: int) -> dict:
    """`config` of the JSON line -- the same dict on the GPU arm and on the reference arm."""
    return {"workload": f"UR5 {B_TRAJ} trajectories x {N_STEPS} steps per GPU, quintic joint_trajectory + "
                        "inverse_dynamics_trajectory (BASELINE.json configs[2])",
            "robot": ROBOT, "trajectories_per_gpu": B_TRAJ, "steps_per_trajectory": N_STEPS,
            "points_per_gpu": B_TRAJ * N_STEPS, "outputs": "float32 torques (B, N, 6)",
            "l2": "256 MiB buffer written between timed steps (outside the event pairs); "
                  "each step also writes 240 MB > 126 MB L2",
            "sharding": f"contiguous trajectory ranges, {world} rank(s), no data-path collective"}


class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled for the whole run -- through NVML (the library
    behind nvidia-smi; every 20 ms) when `pynvml` imports, else `nvidia-smi -lms 100` (which delivers
    about two samples a second with these fields).  `window(t0, t1)` summarises the samples that fell
    into one timed region, so that every entry of the JSON line carries its own clock record."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int, pci_bus_id: str | None = None):
        self.rows, self.proc, self.index, self.pci = [], None, index, pci_bus_id
        self.source, self._stop, self._thread = None, threading.Event(), None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = (pynvml.nvmlDeviceGetHandleByPciBusId(self.pci.encode() if isinstance(self.pci, str) else self.pci)
                 if self.pci else pynvml.nvmlDeviceGetHandleByIndex(self.index))
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons",
                              getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons", None))
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons(h)

            def loop():
                while not self._stop.is_set():
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        bits = int(reasons(h))
                        self.rows.append((time.time(), sm, mx, [n for n, b in self.BITS.items() if bits & b]))
                    except Exception:
                        pass
                    self._stop.wait(0.02)

            self.source = "nvml (20 ms)"
            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
            return self
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 100"
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.rows.append((time.time(), float(r[0]), float(r[1]),
                                  [n for n, v in zip(self.NAMES, r[3:7]) if v.lower().startswith("active")]))
            except (ValueError, IndexError):
                continue

    def _summary(self, rows) -> dict:
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        sm = [r[1] for r in rows]
        reasons = sorted({n for r in rows for n in r[3]})
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(r[2] for r in rows) if rows else None,
                "samples": len(sm), "reasons": reasons, "source": self.source}

    def window(self, t0: float, t1: float) -> dict:
        # (an nvidia-smi sample describes the ~100 ms before it arrived)
        return self._summary([x for x in list(self.rows) if t0 <= x[0] <= t1 + (0.12 if self.proc else 0.0)])

    def stop(self) -> dict:
        time.sleep(0.05 if self._thread else 0.15)
        self._stop.set()
        if self.proc is not None:
            self.proc.terminate()
        return self._summary(list(self.rows))


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


def cpu_baseline_block(rate: float, cores: int, sample: str) -> dict:
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            # the port is compiled C; the reference itself is interpreted Python + NumPy
            "python_reference_points_per_s_per_core": PY_REFERENCE_POINTS_PER_S_PER_CORE,
            "python_reference_source": "BASELINE.md 2: ManipulatorDynamics.inverse_dynamics on the UR5, 61 ms per "
                                       "call on one core of the build container (unmodified reference, float64)",
            "port_points_per_s_per_core": rate / max(1, cores),
            "port_over_python_reference_per_core": rate / max(1, cores) / PY_REFERENCE_POINTS_PER_S_PER_CORE}


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
        "config": headline_config(_env_int("WORLD_SIZE", args.gpus)),
        "arm": "cpu reference algorithm, sampled",
        "cpu_baseline": cpu_baseline_block(rate, cores, desc),
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
    from manipulapy_b200.sharding import PeerRows, gather_rows_pipelined, shard_bounds

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
        import datetime

        # (a rank that parts ways with the others must fail the run quickly, not hold N GPUs for ten minutes)
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=args.pg_timeout))
    ops = _native.ops()
    rb = load_robot(ROBOT, device=dev)
    planner = rb.planner()
    handle = rb.dynamics.robot.handle
    jl = planner._jl
    hbm_peak, peak_src = _measured_peaks()
    g = [0.0, 0.0, -9.81]

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(*xs):
        t = torch.tensor(xs, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    def timed(fn, steps, warmup):
        """seconds summed over `steps` launches of fn (CUDA events on the launching stream, L2 flushed
        between launches outside the event pairs), barrier + synchronize on both sides"""
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
        return sum(a.elapsed_time(b) for a, b in evs) / 1e3

    pr_ = torch.cuda.get_device_properties(dev)
    sampler = ClockSampler(local, f"{pr_.pci_domain_id:08x}:{pr_.pci_bus_id:02x}:{pr_.pci_device_id:02x}.0").start() \
        if rank == 0 else None

    def measure(fn, warmup=3, min_seconds=0.35, min_iters=5, max_iters=400):
        """-> (seconds per launch, launches, clocks): repeats the timed launch until nvidia-smi has had
        time for a few samples, so that every entry carries its own clock record"""
        t0 = time.time()
        sec, n = timed(fn, min_iters, warmup), min_iters
        while time.time() - t0 < min_seconds and n < max_iters:
            k = min(max_iters - n, max(min_iters, n))
            sec += timed(fn, k, 0)
            n += k
        clocks = sampler.window(t0, time.time()) if sampler else None
        return sec / n, n, clocks

    def roof(bytes_per_unit, units, sec, binding, flop_per_unit=None, fp64_peak=None):
        gbs = bytes_per_unit * units / sec / 1e9
        r = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
             "binding": binding, "algorithmic_bytes_per_unit": bytes_per_unit, "peak_source": peak_src}
        if flop_per_unit and fp64_peak:
            tf = flop_per_unit * units / sec / 1e12
            r["fp64"] = {"achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                         "flop_per_unit_textbook": flop_per_unit}
        return r

    # ---- headline: this rank's contiguous shard of the global (world * 4096)-trajectory batch ----
    lo, hi = shard_range(world * B_TRAJ, world, rank)
    rng = np.random.default_rng(3)
    start_all = rng.uniform(-np.pi, np.pi, (world * B_TRAJ, 6))
    end_all = rng.uniform(-np.pi, np.pi, (world * B_TRAJ, 6))
    s_host, e_host = start_all[lo:hi].copy(), end_all[lo:hi].copy()
    s, e = torch.from_numpy(s_host).to(dev), torch.from_numpy(e_host).to(dev)
    B = hi - lo
    P = B * N_STEPS

    if args.mode == "fused":
        def step():
            return ops.trajectory_inverse_dynamics(handle, s, e, False, TF, N_STEPS, METHOD, jl, g, None, None, False)[0]
        launches_per_step, dom_kernel, dom_bytes = 2, "traj_rnea_kernel<6,false>", BYTES_FUSED  # + time-scaling table kernel
    else:
        def step():
            pos, vel, acc = ops.joint_trajectory(s, e, False, TF, N_STEPS, METHOD, jl)
            return ops.inverse_dynamics(handle, pos.view(-1, 6), vel.view(-1, 6), acc.view(-1, 6), g, None, None, None, True)
        launches_per_step, dom_kernel, dom_bytes = 3, "rnea_kernel<6,false>", BYTES_RNEA  # table + trajectory + RNEA

    t_head0 = time.time()
    t_dev = timed(step, args.steps, args.warmup)

    # dominant kernel alone (two_kernel mode: the RNEA launch; fused: the step is that kernel)
    if args.mode == "fused":
        t_dom = t_dev
    else:
        pos, vel, acc = ops.joint_trajectory(s, e, False, TF, N_STEPS, METHOD, jl)
        t_dom = timed(lambda: ops.inverse_dynamics(handle, pos.view(-1, 6), vel.view(-1, 6), acc.view(-1, 6), g,
                                                   None, None, None, True), args.steps, 1)
        del pos, vel, acc

    # PCIe device -> host rate of this box (pinned memory), the ceiling of the e2e number ...
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
    # the per-GPU rate drops (8 GPUs: 57 -> 8-18 GB/s) and differs between GPUs.  Measured over a fixed
    # WINDOW in which every rank copies continuously (64 MiB pieces, 0.3 s): with a fixed amount per rank
    # instead, the ranks on the faster uplink finish first and the others then see less contention than a
    # sustained load gives them (round 2 until now: 11.8 GB/s reported for GPUs 0-3 where the sustained
    # share is ~8.4 GB/s, and a ceiling 12 % too high).
    piece = 64 << 20
    barrier()
    t0 = time.perf_counter()
    done_bytes, t_last, k = 0, t0, 0
    while True:
        lo_ = (k % 4) * piece
        pin[lo_:lo_ + piece].copy_(src[lo_:lo_ + piece], non_blocking=True)
        torch.cuda.synchronize()
        now = time.perf_counter()
        if now - t0 > 0.3:
            break
        done_bytes, t_last, k = done_bytes + piece, now, k + 1
    my_rate = done_bytes / max(1e-9, t_last - t0) / 1e9
    # (keep copying until every rank has left its window, so that nobody's last pieces run uncontended)
    for _ in range(2):
        pin[:piece].copy_(src[:piece], non_blocking=True)
    torch.cuda.synchronize()
    rates = torch.zeros(world, dtype=torch.float64, device=dev)
    rates[rank] = my_rate
    if world > 1:
        dist.all_reduce(rates)
    d2h_rates = [float(x) for x in rates.cpu()]
    del pin, src

    # end to end through the public host API: NumPy endpoints in, NumPy float32 torques out.  The
    # results are host-destined, so at N > 1 the global batch is split in proportion to each rank's
    # measured share of the PCIe uplinks (equal shards pin the step to the slowest link).
    # The copy-rate probe is the starting point; the split is then re-balanced five times from each rank's own
    # measured end-to-end rate (its step time with every other rank running: the uplink shares shift with
    # who is copying when), all outside the timed region.
    e2e_weights = [max(x, 0.02 * max(d2h_rates)) for x in d2h_rates]
    e2e_rebalance = []
    out = None
    E2E_BALANCE_ROUNDS = 5
    for balance_round in range(E2E_BALANCE_ROUNDS + 1 if world > 1 else 1):
        e2e_bounds = shard_bounds(world * B_TRAJ, world, e2e_weights if world > 1 else None)
        elo, ehi = e2e_bounds[rank], e2e_bounds[rank + 1]
        se_host, ee_host = start_all[elo:ehi].copy(), end_all[elo:ehi].copy()

        def e2e_step():
            return planner.trajectory_inverse_dynamics(se_host, ee_host, TF, N_STEPS, METHOD)

        # warm-up in the steady-state pattern of the timed loop (the caller holds the previous result
        # while the next one is produced, so two pinned result buffers are in rotation; the first
        # use of each is a ~100 ms cudaHostAlloc that torch's host allocator then caches; a caller that
        # keeps its results passes out= instead, see OptimizedTrajectoryPlanning.trajectory_inverse_dynamics)
        out = None
        for _ in range(max(3, args.warmup)):
            out = e2e_step()
        if world > 1 and balance_round < E2E_BALANCE_ROUNDS:
            barrier()
            t0 = time.perf_counter()
            for _ in range(6):
                out = e2e_step()
            mine = torch.zeros(world, dtype=torch.float64, device=dev)
            mine[rank] = (ehi - elo) / (time.perf_counter() - t0)
            dist.all_reduce(mine)
            e2e_weights = [float(x) for x in mine.cpu()]
            e2e_weights = [max(x, 0.02 * max(e2e_weights)) for x in e2e_weights]  # (never starve a rank completely)
            e2e_rebalance.append([b_ - a_ for a_, b_ in zip(e2e_bounds, e2e_bounds[1:])])
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = e2e_step()
    barrier()
    t_e2e = time.perf_counter() - t0
    assert out.shape == (ehi - elo, N_STEPS, 6) and out.dtype == np.float32
    del out
    clocks = sampler.window(t_head0, time.time()) if sampler else None

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
    fp64_instr_rot = fma_rate(3)
    fp64_peak_tf = 2 * fp64_instr_peak / 1e12

    # write-only HBM ceiling (hand-written 16-byte stores over 2 GiB): what a pure row-writing kernel can reach
    wbuf = torch.empty(2 << 30, dtype=torch.uint8, device=dev)
    ops.store_peak(wbuf, 0, 148 * 32)
    torch.cuda.synchronize()
    store_peak = 0.0
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.store_peak(wbuf, 0, 148 * 32)
        b.record()
        torch.cuda.synchronize()
        store_peak = max(store_peak, wbuf.numel() / (a.elapsed_time(b) / 1e3) / 1e9)
    del wbuf

    configs = []

    # ---- cfg 4 (second metric of BASELINE.json): forward-dynamics rollout steps/s -----------------
    # iiwa14, 65,536 shooting trajectories x 1000 Euler steps per GPU, float32 torque rows resident in HBM
    t_fd, fd_steps, t_fd_small, fd_clocks = 0.0, 0, 0.0, None
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
        tw0 = time.time()
        t_fd = timed(lambda: ops.forward_dynamics_trajectory(h7, th0, dth0, taum, g, None, 1e-3, 1, jl7), 3, 1)
        fd_steps = Bf * (Nf - 1) * 3
        # one GPU's share of the same 65,536 rollouts on 8 GPUs (strong scaling): each step split across warps
        Bs = Bf // 8
        t_fd_small = timed(lambda: ops.forward_dynamics_trajectory(h7, th0[:Bs], dth0[:Bs], taum[:Bs], g, None, 1e-3, 1,
                                                                   jl7), 3, 1) / 3
        fd_clocks = sampler.window(tw0, time.time()) if sampler else None
        del taum, th0, dth0

    # ---- the other single-GPU configs (N = 1 only: they do not shard differently from the headline) ----
    if world == 1 and not args.no_configs:
        gen = torch.Generator(device=dev).manual_seed(2)

        # cfg 1: the reference's own case -- one UR5 trajectory, N = 1000, two calls
        r1 = np.random.default_rng(1)
        s1, e1 = r1.uniform(-1, 1, 6), r1.uniform(-1, 1, 6)
        s1d, e1d = torch.from_numpy(s1).to(dev).reshape(1, 6), torch.from_numpy(e1).to(dev).reshape(1, 6)

        def cfg1_dev():
            p_, v_, a_ = ops.joint_trajectory(s1d, e1d, True, TF, 1000, METHOD, jl)
            return ops.inverse_dynamics(handle, p_.view(-1, 6), v_.view(-1, 6), a_.view(-1, 6), g, None, None, None, True)

        sec, n, ck = measure(cfg1_dev)

        def cfg1_host():
            tr = planner.joint_trajectory(s1, e1, TF, 1000, METHOD)
            return planner.inverse_dynamics_trajectory(tr["positions"], tr["velocities"], tr["accelerations"])

        for _ in range(5):
            cfg1_host()
        t0 = time.perf_counter()
        for _ in range(50):
            cfg1_host()
        host_s = (time.perf_counter() - t0) / 50
        configs.append({
            "name": "cfg1_ur5_joint_trajectory_plus_inverse_dynamics_trajectory_N1000", "baseline_config": 0,
            "ms": sec * 1e3, "value": 1000 / sec, "unit": "points/s", "launches_timed": n, "gpu_launches_per_call": 2,
            "e2e": {"ms": host_s * 1e3, "value": 1000 / host_s, "unit": "points/s",
                    "api": "planner.joint_trajectory + planner.inverse_dynamics_trajectory, NumPy in / NumPy out"},
            "roofline": roof(72 + 96, 1000, sec, "launch latency: 1000 points occupy 8 of 148 SMs for two ~5 us launches"),
            "clocks": ck,
            "reference_cpu_s": "0.24 s + 61-69 s on one core (BASELINE.md 2)"})

        # cfg 2: FK + space Jacobian over 1 M random configurations; mass matrix over the same
        for name in ("iiwa14", "panda"):
            rbk = load_robot(name, device=dev)
            hk, n_ = rbk.dynamics.robot.handle, rbk.num_joints
            lo_ = torch.from_numpy(rbk.joint_limits[:, 0]).to(dev)
            hi_ = torch.from_numpy(rbk.joint_limits[:, 1]).to(dev)
            Pk = 1_000_000
            th = lo_ + (hi_ - lo_) * torch.rand(Pk, n_, dtype=torch.float64, device=dev, generator=gen)
            sec, n, ck = measure(lambda: ops.fk_jacobian(hk, th, True, True))
            bpu = 8 * n_ + 128 + 48 * n_
            configs.append({
                "name": f"cfg2_fk_plus_space_jacobian_1M_{name}", "baseline_config": 1, "dof": n_,
                "ms": sec * 1e3, "value": Pk / sec, "unit": "configs/s", "launches_timed": n,
                "roofline": roof(bpu, Pk, sec, "hbm", 170 * n_, fp64_peak_tf), "clocks": ck,
                "outputs": "float64 (P, 4, 4) poses + (P, 6, n) Jacobians",
                "l2": "256 MiB flush between launches; outputs 464-512 MB > L2"})
            sec, n, ck = measure(lambda: ops.mass_matrix(hk, th))
            bpu = 8 * n_ + 8 * n_ * n_
            configs.append({
                "name": f"mass_matrix_crba_1M_{name}", "baseline_config": 1, "dof": n_,
                "ms": sec * 1e3, "value": Pk / sec, "unit": "configs/s", "launches_timed": n,
                "roofline": roof(bpu, Pk, sec, "hbm", 64 * n_ + 100 * n_ + 53 * n_ * (n_ + 1) // 2, fp64_peak_tf),
                "clocks": ck, "outputs": "float64 (P, n, n)"})
            del th

        # the trajectory kernel alone (rows of positions, velocities, accelerations: 72 B / point)
        sec, n, ck = measure(lambda: ops.joint_trajectory(s, e, False, TF, N_STEPS, METHOD, jl))
        rt = roof(BYTES_TRAJ, P, sec, "hbm (write only)")
        rt["write_only_peak_gbs"] = store_peak
        rt["frac_of_write_only_peak"] = rt["achieved"] / store_peak
        configs.append({"name": "joint_trajectory_rows_ur5_4096x2441", "baseline_config": 2, "ms": sec * 1e3,
                        "value": P / sec, "unit": "points/s", "launches_timed": n, "roofline": rt, "clocks": ck,
                        "outputs": "3 x float32 (B, N, 6)"})

        # the collision / limit hook of joint_trajectory (>= 99 % of cfg 1's wall time in the
        # reference: a host loop of link_fk + AABB tests per row).  Synthetic per-link point sets (the reference's
        # hulls come from meshes, which nothing here loads): link FK of the whole link tree and the self-collision
        # flags over 1 M random UR5 configurations, and cfg 1's trajectory with the hook applied.
        hr = np.random.default_rng(11)
        link_names = [str(x) for x in rb.links["link_names"]]
        hulls = {nm: hr.uniform(-0.03, 0.03, 3) + hr.uniform(-1, 1, (16, 3)) * hr.uniform(0.05, 0.10, 3)
                 for nm in link_names[3:9]}  # the six moving links
        checker = rb.collision_checker(hulls)
        Pk, L = 1_000_000, len(link_names)
        lo_ = torch.from_numpy(rb.joint_limits[:, 0]).to(dev)
        hi_ = torch.from_numpy(rb.joint_limits[:, 1]).to(dev)
        th = lo_ + (hi_ - lo_) * torch.rand(Pk, 6, dtype=torch.float64, device=dev, generator=gen)
        sec, n, ck = measure(lambda: ops.link_fk_batch(handle, checker._model_host, checker._model_dev, th, L))
        configs.append({"name": "link_fk_batch_1M_ur5", "survey_row": "8f-1", "links": L, "ms": sec * 1e3,
                        "value": Pk / sec, "unit": "configs/s", "launches_timed": n,
                        "roofline": roof(48 + 128 * L, Pk, sec, "hbm"), "clocks": ck,
                        "outputs": "float64 (P, L, 4, 4) poses of every link"})
        sec, n, ck = measure(lambda: ops.self_collision(handle, checker._model_host, checker._model_dev, th))
        flags = ops.self_collision(handle, checker._model_host, checker._model_dev, th)
        configs.append({"name": "self_collision_aabb_1M_ur5", "survey_row": "8f-1", "hulls": len(hulls),
                        "points_per_hull": 16, "ms": sec * 1e3, "value": Pk / sec, "unit": "configs/s",
                        "launches_timed": n, "colliding_fraction": float(flags.float().mean()),
                        "roofline": roof(48 + 1, Pk, sec, "fp64 pipe (6 hulls x 16 points x 9 FMA + 15 pair tests per row)",
                                         6 * 16 * 18 + 170 * 6, fp64_peak_tf), "clocks": ck,
                        "outputs": "uint8 (P,) flags"})
        # cfg 1's call with the hook on, from a COLLIDING start configuration (so that rows are actually nudged)
        hit = torch.nonzero(flags)
        s1c = th[int(hit[0])].cpu().numpy() if hit.numel() else s1
        del th, flags
        hook_planner = rb.planner()
        hook_planner.attach_collision_checker(checker)
        for _ in range(5):
            hook_planner.joint_trajectory(s1c, e1, TF, 1000, METHOD)
        t0 = time.perf_counter()
        for _ in range(50):
            tr_hook = hook_planner.joint_trajectory(s1c, e1, TF, 1000, METHOD)
        hook_s = (time.perf_counter() - t0) / 50
        t0 = time.perf_counter()
        for _ in range(50):
            planner.joint_trajectory(s1c, e1, TF, 1000, METHOD)
        plain_s = (time.perf_counter() - t0) / 50
        configs.append({"name": "cfg1_joint_trajectory_with_collision_hook_N1000", "survey_row": "8f-1", "baseline_config": 0,
                        "ms": hook_s * 1e3, "value": 1000 / hook_s, "unit": "points/s",
                        "ms_without_hook": plain_s * 1e3,
                        "api": "planner.joint_trajectory with a CollisionChecker attached, NumPy in / NumPy out (host time)",
                        "rows_nudged": int((tr_hook["positions"] != planner.joint_trajectory(s1c, e1, TF, 1000, METHOD)["positions"])
                                           .any(axis=1).sum()),
                        "reference_cpu_s": "the host hook is >= 99 % of the reference's joint_trajectory wall time (SURVEY 8a1)"})

    # ---- cfg 5: 1e6 .. 1e9 points, strong scaling, gathered on rank 0 inside the timed region ----
    sweep = []
    if not args.no_sweep:
        side = torch.cuda.Stream(device=dev)
        for Ptot in (10 ** 6, 10 ** 7, 10 ** 8, 10 ** 9):
            if Ptot > args.sweep_max:
                continue
            Bt = -(-Ptot // N_STEPS)
            # (even shard boundaries: 2441 x 6 x 4 bytes per trajectory is 8 mod 16, and a shard of the shared
            # result buffer that starts on a 16-byte boundary is written with full-width vector stores)
            blo, bhi = shard_range(Bt, world, rank, align=2)
            Bl = bhi - blo
            gen5 = torch.Generator(device=dev).manual_seed(5)
            # (every rank draws the same global endpoint table and keeps its rows: cheap, 48 B per 58 KB of output)
            ends = (torch.rand(2, Bt, 6, dtype=torch.float64, device=dev, generator=gen5) * 2 - 1) * np.pi
            s5, e5 = ends[0, blo:bhi].contiguous(), ends[1, blo:bhi].contiguous()
            # launches per timing: the same on every rank (a function of the sizes only), enough of them that
            # the entry's clock record has several NVML samples
            est = Bt * N_STEPS / world / 1.7e10 + 1.2e-4
            steps5 = int(min(400, max(3, -(-0.12 // est))))
            tw0 = time.time()

            def launch_into(dest, a_=0, b_=None):
                b_ = Bl if b_ is None else b_
                if b_ > a_:
                    ops.trajectory_inverse_dynamics(handle, s5[a_:b_], e5[a_:b_], False, TF, N_STEPS, METHOD, jl, g, None,
                                                    None, False, False, dest)

            entry = {"points": Bt * N_STEPS, "trajectories": Bt, "trajectories_this_rank": Bl,
                     "result_bytes": Bt * N_STEPS * 24}
            # (a) compute only: every rank its shard into its own HBM
            local_out = torch.empty((max(Bl, 1), N_STEPS, 6), dtype=torch.float32, device=dev)
            t_c = timed(lambda: launch_into(local_out[:Bl]), steps5, 2) / steps5
            (t_c,) = reduce_max(t_c)
            entry["compute_ms"] = t_c * 1e3
            entry["points_per_s_compute_only"] = Bt * N_STEPS / t_c
            if world > 1:
                # (b) fused: the kernel stores its rows into rank 0's buffer over NVLink (PeerRows).  Whether the
                # mapping and a first launch work is decided by all ranks together: a rank must never part ways
                pr, perr = None, None
                try:
                    pr = PeerRows(Bt, (N_STEPS, 6), torch.float32, dev, dst=0)
                except RuntimeError as ex:  # (raised on every rank: agreed inside PeerRows)
                    perr = f"{type(ex).__name__}: {ex}"[:200]
                if pr is not None:
                    try:
                        launch_into(pr.rows())
                    except Exception as ex:
                        perr = f"{type(ex).__name__}: {ex}"[:200]
                    okf = torch.tensor([0 if perr else 1], dtype=torch.int32, device=dev)
                    dist.all_reduce(okf, op=dist.ReduceOp.MIN)
                    if int(okf) == 0:
                        perr = perr or "the first launch into the peer-mapped buffer failed on another rank"
                if pr is not None and perr is None:
                    def fused():
                        launch_into(pr.rows())
                        pr.commit()

                    t_p = timed(fused, steps5, 2) / steps5
                    (t_p,) = reduce_max(t_p)
                    entry["peer_store_ms"] = t_p * 1e3
                    entry["gathered_points_per_s"] = Bt * N_STEPS / t_p
                    entry["gather"] = "kernel stores into rank 0's peer-mapped buffer over NVLink (fused)"
                    if rank == 0 and Ptot <= 10 ** 7:
                        # parity of the gathered result with a local recomputation of the whole batch
                        ref = torch.empty((Bt, N_STEPS, 6), dtype=torch.float32, device=dev)
                        ops.trajectory_inverse_dynamics(handle, ends[0], ends[1], False, TF, N_STEPS, METHOD, jl, g, None,
                                                        None, False, False, ref)
                        entry["gathered_equals_single_gpu_bits"] = bool(torch.equal(ref, pr.full))
                        del ref
                if pr is not None:
                    pr.close()
                    del pr
                if perr:
                    entry["peer_store_error"] = perr
                # (c) NCCL baseline: chunked isend / irecv on a second stream while the next chunk computes
                full = torch.empty((Bt, N_STEPS, 6), dtype=torch.float32, device=dev) if rank == 0 else None

                def nccl_pipe(chunks=8, compute=True):
                    def lf(a_, b_, dest):
                        if compute:
                            launch_into(dest, a_ - blo, b_ - blo)
                    return gather_rows_pipelined(lf, Bt, (N_STEPS, 6), torch.float32, dev, dst=0, chunks=chunks,
                                                 out=full, side=side, align=2)

                t_n = timed(lambda: nccl_pipe(8), steps5, 2) / steps5
                t_g = timed(lambda: nccl_pipe(1, False), steps5, 1) / steps5  # the transfer alone
                t_n, t_g = reduce_max(t_n, t_g)
                entry["nccl_pipelined_ms"], entry["nccl_transfer_only_ms"] = t_n * 1e3, t_g * 1e3
                entry["nccl_gathered_points_per_s"] = Bt * N_STEPS / t_n
                entry["nccl_overlap_fraction"] = max(0.0, min(1.0, (t_c + t_g - t_n) / max(1e-12, min(t_c, t_g))))
                entry["nvlink_ingest_gbs_rank0"] = (Bt - (shard_bounds(Bt, world, None, 2)[1])) * N_STEPS * 24 / t_g / 1e9
                if "peer_store_ms" in entry:
                    entry["nvlink_ingest_gbs_rank0_peer_store"] = ((Bt - shard_bounds(Bt, world, None, 2)[1]) * N_STEPS * 24
                                                                   / (entry["peer_store_ms"] / 1e3) / 1e9)
                if "gathered_points_per_s" not in entry:
                    entry["gathered_points_per_s"] = entry["nccl_gathered_points_per_s"]
                    entry["gather"] = "NCCL isend / irecv, 8 chunks, second stream"
                elif entry["nccl_gathered_points_per_s"] > entry["gathered_points_per_s"]:
                    # (ingest-bound sizes at N >= 4: the copy-engine path moves bytes a few per cent faster than SM
                    # stores over the link; both numbers stay in the entry)
                    entry["gathered_points_per_s"] = entry["nccl_gathered_points_per_s"]
                    entry["gather"] = ("NCCL isend / irecv, 8 chunks, second stream (faster than the fused peer store at "
                                       "this size: %.3f against %.3f ms)" % (t_n * 1e3, entry["peer_store_ms"]))
                del full
            else:
                entry["gathered_points_per_s"] = entry["points_per_s_compute_only"]
                entry["gather"] = "single rank: the result is already on rank 0"
            entry["clocks"] = sampler.window(tw0, time.time()) if sampler else None
            entry["roofline"] = roof(BYTES_FUSED, Bt * N_STEPS / world, t_c, "fp64_pipe", FLOP_EXECUTED_PER_POINT["fused"],
                                     fp64_peak_tf)
            sweep.append(entry)
            del local_out, ends, s5, e5

    t_dev, t_dom, t_e2e, t_fd = reduce_max(t_dev, t_dom, t_e2e, t_fd)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_pts = world * P * args.steps
    value = total_pts / t_dev
    dom_s = t_dom / args.steps
    achieved_gbs = dom_bytes * P / dom_s / 1e9
    achieved_tf = FLOP_EXECUTED_PER_POINT[args.mode] * P / dom_s / 1e12
    instr_rate = FP64_INSTR_PER_POINT[args.mode] * P / dom_s
    cpu_rate, cores, cpu_pts, _ = cpu_reference_rate(args.cpu_sample_traj, 1, 0) if args.gpus == 1 and not args.no_cpu else (None, None, None, None)
    cfg = headline_config(world)
    d2h_min, d2h_sum = min(d2h_rates), sum(d2h_rates)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg, "kernel_mode": args.mode,
        "clocks": clocks,
        "e2e": {"value": world * B_TRAJ * N_STEPS * args.steps / t_e2e, "unit": UNIT,
                "h2d_bytes_per_step": int(2 * (ehi - elo) * 6 * 8), "d2h_bytes_per_step": int((ehi - elo) * N_STEPS * 6 * 4),
                "api": "OptimizedTrajectoryPlanning.trajectory_inverse_dynamics",
                "pcie_d2h_gbs_measured": d2h_gbs, "pcie_d2h_gbs_all_ranks_together": d2h_rates,
                "pcie_d2h_gbs_all_ranks_together_min": d2h_min, "pcie_d2h_gbs_all_ranks_together_sum": d2h_sum,
                "host_cpus_bound_to": numa_cpus,
                "sharding": ("trajectories split in proportion to each rank's measured end-to-end rate (start: the "
                             "concurrent device->host copy rates; re-balanced five times outside the timed region): "
                             + str([b_ - a_ for a_, b_ in zip(e2e_bounds, e2e_bounds[1:])])) if world > 1 else "one rank",
                "sharding_rebalance_history": e2e_rebalance,
                # ceilings of the e2e number: every result byte crosses PCIe; with weighted shards the
                # aggregate rate counts, with equal shards the slowest rank's
                "pcie_bound_points_per_s": d2h_sum * 1e9 / (6 * 4),
                "pcie_bound_points_per_s_equal_shards": world * d2h_min * 1e9 / (6 * 4)},
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
                              "fp64_instr_per_point": FP64_INSTR_PER_POINT[args.mode],
                              "instr_peak_per_s": fp64_instr_peak,
                              # sustained rates of the pipe for the operand patterns the kernel is made of
                              # (register-file bandwidth, not the pipe, bounds these: DESIGN.md 4)
                              "instr_rate_three_register_operands_per_s": fp64_instr_3reg,
                              "instr_rate_planar_rotation_pattern_per_s": fp64_instr_rot,
                              "peak_source": "mpk_fma_peak measured in this run (mode 0: shared operands = pipe peak; "
                                             "mode 1: three distinct register operands; mode 3: DMUL + DFMA rotation)"}},
        "write_only_hbm_peak_gbs": store_peak,
    }
    if t_fd > 0:
        fd_entry = {
            "metric": "fd_rollout_steps_per_s", "value": world * fd_steps / t_fd, "unit": "steps/s",
            "ms_per_launch": t_fd / 3 * 1e3, "gpu_launches": 3,
            "ms_per_launch_8192_rollouts": t_fd_small * 1e3,  # rank 0's time (not reduced over ranks)
            "strong_scaling_8_gpu_speedup_estimate": (t_fd / 3) / t_fd_small,
            "clocks": fd_clocks,
            "roofline": roof(28 + 84, FD_ROLLOUTS * (FD_STEPS - 1), t_fd / 3, "fp64_pipe + sequential latency", 5300, fp64_peak_tf),
            "config": {"workload": f"iiwa14 {FD_ROLLOUTS} rollouts x {FD_STEPS} Euler steps per GPU, CRBA mass matrix + "
                                   "LDL^T solve per step (BASELINE.json configs[3])", "dt": 1e-3, "intRes": 1,
                       "theta0": "U(0.5 lo, 0.5 hi)", "dtheta0": "U(-0.5, 0.5)",
                       "taumat": "float32 (B, N, 7) resident in HBM: gravity compensation at theta0 + per-joint uniform "
                                 "noise of amplitude [4, 4, 2, 2, 0.4, 0.2, 0.08] N m (SURVEY 8d's literal U(-20, 20) makes "
                                 "explicit Euler overflow for a few rollouts, in the reference too; that distribution is "
                                 "covered at full size by test_full_size_cfg4_literal_torques: ~3 % of the rollouts overflow, at the oracle's step)",
                       "outputs": "3 x float32 (B, N, 7)"}}
        line["fd_rollout"] = fd_entry
        configs.append(dict(fd_entry, name="cfg4_iiwa14_forward_dynamics_rollouts_65536x1000", baseline_config=3,
                            ms=t_fd / 3 * 1e3))
    if sweep:
        configs.append({"name": "cfg5_ur5_trajectory_plus_rnea_sweep", "baseline_config": 4, "scaling": "strong",
                        "n_gpus": world, "unit": "points/s",
                        "what": "fused trajectory + inverse dynamics of ceil(P / 2441) trajectories split by shard_range "
                                "over the ranks; float32 torque rows gathered on rank 0 inside the timed region",
                        "sizes": sweep})
    line["configs"] = configs
    if cpu_rate is not None:
        line["cpu_baseline"] = cpu_baseline_block(
            cpu_rate, cores,
            f"{args.cpu_sample_traj} of {B_TRAJ} trajectories x {N_STEPS} steps ({cpu_pts} points), "
            "oracle/oracle.c literal port of the reference algorithm, all host threads")
    if sampler:
        sampler.stop()
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
    ap.add_argument("--no-configs", action="store_true", help="skip the single-GPU entries of `configs`")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 1e6 .. 1e9-point strong-scaling sweep")
    ap.add_argument("--pg-timeout", type=int, default=180, help="seconds before a stuck collective aborts the run")
    ap.add_argument("--sweep-max", type=float, default=1e9, help="largest point count of the sweep")
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
