"""``OptimizedTrajectoryPlanning`` -- drop-in mirror of ``ManipulaPy.path_planning``
(planning/trajectory_planning.py:116-817) for the batched trajectory-and-dynamics hot path.

=================================  =============================================  ============
method                             reference                                      returns
=================================  =============================================  ============
``joint_trajectory``               planning/trajectory.py:103-169, 276-333        dict of 3 x ``(N, n)`` float32
``batch_joint_trajectory``         planning/trajectory.py:335-502                 dict of 3 x ``(B, N, n)`` float32
``inverse_dynamics_trajectory``    planning/trajectory_dynamics.py:31-90,308-380  ``(P, n)`` float32, clipped
``forward_dynamics_trajectory``    planning/trajectory_dynamics.py:382-423,580-708 dict of 3 x ``(N, n)`` float32
``trajectory_inverse_dynamics``    (extension) the first and third fused          ``(B, N, n)`` float32
=================================  =============================================  ============

Semantics reproduced from the reference's CPU path (the parity target): float32 outputs,
float32-rounded joint / torque limits (:218-223), position clip after generation, torque
row cast to float32 then clipped, semi-implicit Euler with ``intRes`` sub-steps, ``theta``
clip without velocity reset, row 0 = initial state, ``IndexError`` for an empty ``taumat``.
There is no CPU routing: every call runs the CUDA kernels (``use_cuda=False`` raises).
The collision / potential-field post-processing hook of ``joint_trajectory``
(planning/trajectory.py:311-317, planning/collision_host.py:40-88) is applied when a
``collision_checker`` (``manipulapy_b200.CollisionChecker``) and a ``potential_field`` are attached:
the reference builds its checker from the URDF's meshes, this package from link data and hull points
(``planner.attach_collision_checker(hulls)``), since mesh parsing is outside the hot path.
"""

from __future__ import annotations

import logging
import time
from typing import Any, Dict, Optional

import numpy as np
import torch

from . import _host, _native
from .cuda_kernels import KERNEL_REGISTRY

logger = logging.getLogger("manipulapy_b200.path_planning")

_KERNEL_TYPES = ("auto", "standard", "vectorized", "memory_optimized", "warp_optimized",
                 "cache_friendly", "auto_tune")


class OptimizedTrajectoryPlanning:
    def __init__(self, serial_manipulator, urdf_path, dynamics, joint_limits, torque_limits=None, *,
                 use_cuda: Optional[bool] = None, cuda_threshold: int = 10,
                 memory_pool_size_mb: Optional[int] = None, enable_profiling: bool = False,
                 auto_optimize: bool = True, kernel_type: str = "auto", target_speedup: float = 40.0,
                 device: Optional[Any] = None) -> None:
        if use_cuda is False:
            raise RuntimeError(
                "use_cuda=False requested, but manipulapy_b200 has no CPU path; use the reference "
                "ManipulaPy planner for CPU execution")
        if not torch.cuda.is_available():
            raise RuntimeError(
                "use_cuda=True requested but no GPU-capable backend with CUDA is active."
                if use_cuda else "manipulapy_b200 needs a CUDA device (no CPU fallback)")
        self.serial_manipulator = serial_manipulator
        self.dynamics = dynamics
        self.urdf_path = urdf_path
        # float32-rounded limits, as the reference stores them (trajectory_planning.py:218-223)
        self.joint_limits = np.asarray(joint_limits, dtype=np.float32)
        self.torque_limits = (
            np.asarray(torque_limits, dtype=np.float32) if torque_limits is not None
            else np.array([[-np.inf, np.inf]] * len(joint_limits), dtype=np.float32))
        self.kernel_type = kernel_type if kernel_type is not None else "auto"
        self.target_speedup = target_speedup
        self.enable_profiling = bool(enable_profiling)
        self.cuda_available = True
        self.cpu_threshold = 0  # everything runs on the GPU
        # the reference builds CollisionChecker(urdf_path) here (trajectory_planning.py:232-238); without its
        # mesh loader the hook starts detached -- attach_collision_checker() switches it on
        self.collision_checker = None
        self.potential_field = None
        if urdf_path is not None:
            logger.warning("urdf_path is given, but manipulapy_b200 does not parse URDF meshes: the collision "
                           "hook of joint_trajectory stays off until attach_collision_checker(hulls) is called")
        self.device = _host.default_device(device if device is not None
                                           else getattr(dynamics, "_device_arg", None))
        p = torch.cuda.get_device_properties(self.device)
        self.gpu_properties = {"multiprocessor_count": p.multi_processor_count,
                               "max_threads_per_block": 1024, "name": p.name}
        self.performance_stats = {
            "gpu_calls": 0, "cpu_calls": 0, "total_gpu_time": 0.0, "total_cpu_time": 0.0,
            "memory_transfers": 0, "kernel_launches": 0, "speedup_achieved": 0.0,
            "best_kernel_used": "none",
        }
        self._jl = _host.limits_tensor(self.joint_limits)
        self._tl = _host.limits_tensor(self.torque_limits)
        # host results of the fused path leave the device in this many pipelined chunks
        self.host_chunks = 8

    # -- bookkeeping ------------------------------------------------------------------------------
    def _should_use_gpu(self, N: int, num_joints: int) -> bool:
        return True

    def _tick(self, t0: float, launches: int = 1, transfers: int = 0, kernel: str = "mpk") -> None:
        s = self.performance_stats
        s["gpu_calls"] += 1
        s["kernel_launches"] += launches
        s["memory_transfers"] += transfers
        s["total_gpu_time"] += time.perf_counter() - t0
        s["best_kernel_used"] = kernel

    def get_performance_stats(self) -> Dict[str, Any]:
        s = dict(self.performance_stats)
        s["avg_gpu_time"] = s["total_gpu_time"] / s["gpu_calls"] if s["gpu_calls"] else 0.0
        s["avg_cpu_time"] = 0.0
        s["gpu_usage_percent"] = 100.0 if s["gpu_calls"] else 0.0
        return s

    def reset_performance_stats(self) -> None:
        for k, v in self.performance_stats.items():
            self.performance_stats[k] = "none" if isinstance(v, str) else type(v)(0)

    def _check_kernel_type(self, kernel_type: Optional[str]) -> str:
        kt = self.kernel_type if kernel_type is None else kernel_type
        KERNEL_REGISTRY.get(f"trajectory.{kt}")  # KeyError for unknown names, like the reference
        return kt

    def attach_collision_checker(self, checker_or_hulls, links=None, potential_field=None) -> None:
        """Switch on the collision hook of ``joint_trajectory``: a ``CollisionChecker``, or
        ``{link name: (V, 3) hull points}`` together with the robot's link table."""
        from .potential_field import CollisionChecker, PotentialField

        if isinstance(checker_or_hulls, CollisionChecker):
            self.collision_checker = checker_or_hulls
        else:
            if links is None:
                raise ValueError("links (the robot's link table) is required with a hull dictionary")
            self.collision_checker = CollisionChecker(self.dynamics, links, checker_or_hulls, device=self.device)
        self.potential_field = potential_field or PotentialField()

    # -- trajectory generation ------------------------------------------------------------------------
    def joint_trajectory(self, thetastart, thetaend, Tf, N, method, kernel_type=None,
                         enable_monitoring=None) -> Dict[str, Any]:
        t0 = time.perf_counter()
        self._check_kernel_type(kernel_type)
        on_dev = _host.any_device(thetastart, thetaend)
        dev = thetastart.device if _host.is_device_tensor(thetastart) else self.device
        s = _host.to_device(thetastart, dev).reshape(1, -1)
        e = _host.to_device(thetaend, dev).reshape(1, -1)
        pos, vel, acc = _native.ops().joint_trajectory(s, e, True, float(Tf), int(N), int(method), self._jl)
        if self.collision_checker and self.potential_field and int(N) > 0:
            # planning/trajectory.py:316-317: colliding rows are nudged towards thetaend (float32, as cast above)
            self.collision_checker.avoid(pos[0], e.float(), self.potential_field)
        out = {"positions": pos[0], "velocities": vel[0], "accelerations": acc[0]}
        if not on_dev:
            out = {k: _host.to_host(v) for k, v in out.items()}
        self._tick(t0, transfers=0 if on_dev else 5, kernel="trajectory")
        return out

    def batch_joint_trajectory(self, thetastart_batch, thetaend_batch, Tf, N, method,
                               kernel_type=None) -> Dict[str, Any]:
        t0 = time.perf_counter()
        self._check_kernel_type(kernel_type)
        on_dev = _host.any_device(thetastart_batch, thetaend_batch)
        f32 = _input_is_f32(thetastart_batch) and _input_is_f32(thetaend_batch)
        dev = thetastart_batch.device if _host.is_device_tensor(thetastart_batch) else self.device
        s = _host.to_device(thetastart_batch, dev)
        e = _host.to_device(thetaend_batch, dev)
        if s.dim() != 2:
            raise ValueError("thetastart_batch must be (batch_size, num_joints)")
        pos, vel, acc = _native.ops().joint_trajectory(s, e, f32, float(Tf), int(N), int(method), self._jl)
        out = {"positions": pos, "velocities": vel, "accelerations": acc}
        if not on_dev:
            out = {k: _host.to_host(v) for k, v in out.items()}
        self._tick(t0, transfers=0 if on_dev else 5, kernel="trajectory")
        return out

    # -- dynamics over trajectories ------------------------------------------------------------------------
    def inverse_dynamics_trajectory(self, thetalist_trajectory, dthetalist_trajectory,
                                    ddthetalist_trajectory, gravity_vector=None, Ftip=None, precision=None):
        """``precision="float32"`` (extension): float32 arithmetic, 1e-4 relative on torques."""
        t0 = time.perf_counter()
        on_dev = _host.any_device(thetalist_trajectory, dthetalist_trajectory, ddthetalist_trajectory)
        dyn = self.dynamics
        n = dyn.num_joints
        dev = thetalist_trajectory.device if _host.is_device_tensor(thetalist_trajectory) else self.device
        th = _host.to_device(thetalist_trajectory, dev, keep_f32=True)
        shape = tuple(th.shape)
        th = th.reshape(-1, n)
        dth = _host.to_device(dthetalist_trajectory, dev, keep_f32=True).reshape(-1, n)
        ddth = _host.to_device(ddthetalist_trajectory, dev, keep_f32=True).reshape(-1, n)
        th, dth, ddth = _host.promote_rows(th, dth, ddth)
        ftip = None if Ftip is None else _host.vec(Ftip, 6, "Ftip")
        tau = _native.ops().inverse_dynamics(dyn.robot.handle, th, dth, ddth, _host.gravity(gravity_vector),
                                             ftip, None, self._tl, True, _host.is_f32(precision)).reshape(shape)
        out = tau if on_dev else _host.to_host(tau)
        self._tick(t0, transfers=0 if on_dev else 4, kernel="inverse_dynamics")
        return out

    def trajectory_inverse_dynamics(self, thetastart_batch, thetaend_batch, Tf, N, method,
                                    gravity_vector=None, Ftip=None, return_trajectory: bool = False,
                                    precision=None, out=None):
        """``batch_joint_trajectory`` followed by ``inverse_dynamics_trajectory`` in ONE kernel.

        Identical results to the two calls (the trajectory rows are rounded to float32 and
        clipped in registers before the dynamics), without the ``3 x 12 n`` bytes per point
        of HBM round trip.  Returns ``(B, N, n)`` float32 torques (and the trajectory dict
        when ``return_trajectory``).  ``precision="float32"``: the trajectory rows are unchanged
        (bit-exact), the inverse dynamics runs in float32 arithmetic (1e-4 relative on torques).

        ``out`` (extension): caller-owned destination of the torques.  Host inputs: a ``(B, N, n)``
        float32 NumPy array / CPU tensor, ideally pinned, filled chunk by chunk (host results
        otherwise come back in freshly pinned memory, which a caller that keeps them accumulates).
        Device inputs: a contiguous float32 CUDA tensor of ``B * N * n`` elements -- it may be another
        GPU's peer-mapped memory (``sharding.PeerRows``): the kernel then stores straight over
        NVLink and the multi-GPU gather costs no extra pass.
        """
        f32c = _host.is_f32(precision)
        t0 = time.perf_counter()
        on_dev = _host.any_device(thetastart_batch, thetaend_batch)
        a = np.asarray(thetastart_batch) if not on_dev else thetastart_batch
        single = a.ndim == 1
        f32 = single or (_input_is_f32(thetastart_batch) and _input_is_f32(thetaend_batch))
        dev = thetastart_batch.device if _host.is_device_tensor(thetastart_batch) else self.device
        n = self.dynamics.num_joints
        s = _host.to_device(thetastart_batch, dev).reshape(-1, n)
        e = _host.to_device(thetaend_batch, dev).reshape(-1, n)
        ftip = None if Ftip is None else _host.vec(Ftip, 6, "Ftip")
        ops, handle, g = _native.ops(), self.dynamics.robot.handle, _host.gravity(gravity_vector)
        if out is not None and (return_trajectory or (on_dev != _host.is_device_tensor(out))):
            raise ValueError("out= needs return_trajectory=False and must live where the inputs live "
                             "(host array for host inputs, CUDA tensor for device inputs)")
        if not on_dev and not return_trajectory and (s.shape[0] >= 64 or out is not None):
            # host result: pipeline the kernel with the device->host copy, chunk by chunk
            def launch(lo, hi):
                return ops.trajectory_inverse_dynamics(handle, s[lo:hi], e[lo:hi], f32, float(Tf), int(N),
                                                       int(method), self._jl, g, ftip, self._tl, False, f32c)[0]

            B = int(s.shape[0])
            res = _host.chunked_to_host(launch, B, (int(N), n), torch.float32, dev, chunks=self.host_chunks,
                                        min_rows=64 if B >= 64 else 1, out=out)
            self._tick(t0, launches=2 * min(self.host_chunks, B), transfers=2 + min(self.host_chunks, B),
                       kernel="trajectory_inverse_dynamics")
            return res
        tau, pos, vel, acc = ops.trajectory_inverse_dynamics(
            handle, s, e, f32, float(Tf), int(N), int(method), self._jl, g, ftip, self._tl,
            bool(return_trajectory), f32c, out)
        if out is not None:
            self._tick(t0, launches=2, kernel="trajectory_inverse_dynamics")
            return out
        outs = [tau] + ([pos, vel, acc] if return_trajectory else [])
        if single:
            outs = [o[0] for o in outs]
        if not on_dev:
            outs = [_host.to_host(o) for o in outs]
        self._tick(t0, launches=2, transfers=0 if on_dev else 2 + len(outs), kernel="trajectory_inverse_dynamics")
        if return_trajectory:
            return outs[0], {"positions": outs[1], "velocities": outs[2], "accelerations": outs[3]}
        return outs[0]

    def cartesian_trajectory(self, Xstart, Xend, Tf, N, method) -> Dict[str, Any]:
        """Straight-line Cartesian trajectory (planning/trajectory.py:504-594): orientation
        ``Rstart exp(log(Rstart^T Rend) s)``, position ``s pend + (1 - s) pstart`` and the linear
        velocity / acceleration, float32.  Reference call: ``(4, 4)`` poses -> ``(N, 3)`` /
        ``(N, 3, 3)`` arrays; batched extension: ``(B, 4, 4)`` -> ``(B, N, 3)`` / ``(B, N, 3, 3)``."""
        t0 = time.perf_counter()
        N = int(N)
        if N < 0:
            raise ValueError("negative dimensions are not allowed")
        if N == 1:
            raise ZeroDivisionError("float division by zero")  # Tf / (N - 1.0), trajectory.py:525
        on_dev = _host.any_device(Xstart, Xend)
        dev = Xstart.device if _host.is_device_tensor(Xstart) else (
            Xend.device if _host.is_device_tensor(Xend) else self.device)
        xs, xe = _host.to_device(Xstart, dev), _host.to_device(Xend, dev)
        single = xs.dim() == 2
        xs, xe = xs.reshape(-1, 4, 4), xe.reshape(-1, 4, 4)
        if xs.shape != xe.shape:
            raise ValueError("Xstart and Xend must have the same shape")
        pos, vel, acc, ori = _native.ops().cartesian_trajectory(xs, xe, float(Tf), N, int(method))
        outs = [pos, vel, acc, ori]
        if single:
            outs = [o[0] for o in outs]
            if N == 0:
                outs[0] = outs[0].reshape(0)  # the reference's empty positions array is (0,)
        if not on_dev:
            outs = [_host.to_host(o) for o in outs]
        self._tick(t0, transfers=0 if on_dev else 6, kernel="cartesian_trajectory")
        return {"positions": outs[0], "velocities": outs[1], "accelerations": outs[2], "orientations": outs[3]}

    def forward_dynamics_trajectory(self, thetalist, dthetalist, taumat, g, Ftipmat, dt, intRes) -> Dict[str, Any]:
        """Single rollout (reference call: ``thetalist (n,)``, ``taumat (N, n)``, ``Ftipmat (N, 6)``)
        or ``B`` independent rollouts (``(B, n)``, ``(B, N, n)``, ``(B, N, 6)`` or ``None``).

        The state is integrated in float64 whatever the dtype of ``thetalist`` / ``dthetalist``.  (Deviation: the
        reference keeps the state in the INPUT dtype -- trajectory_dynamics.py:619-678 casts back after every
        sub-step -- so a float32 initial state, e.g. ``traj["positions"][0]``, makes it integrate, and evaluate its
        trigonometry, in float32; here such a state is upcast exactly and the rollout equals the reference's on
        the float64 copy of the same state.)"""
        t0 = time.perf_counter()
        on_dev = _host.any_device(thetalist, dthetalist, taumat)
        dyn = self.dynamics
        n = dyn.num_joints
        dev = taumat.device if _host.is_device_tensor(taumat) else self.device
        th0 = _host.to_device(thetalist, dev)
        single = th0.dim() == 1
        th0 = th0.reshape(-1, n)
        B = th0.shape[0]
        dth0 = _host.to_device(dthetalist, dev).reshape(B, n)
        tm = _host.to_device(taumat, dev, keep_f32=True)
        N = int(tm.shape[-2]) if tm.dim() >= 2 else 0
        if N == 0 or tm.numel() == 0:
            # the reference indexes row 0 of an empty result (trajectory_dynamics.py:612-615)
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")
        tm = tm.reshape(B, N, n)
        fm = None
        if Ftipmat is not None:
            fm = _host.to_device(Ftipmat, dev).reshape(B, N, 6)
            if not bool(torch.any(fm != 0)):
                fm = None
        pos, vel, acc = _native.ops().forward_dynamics_trajectory(
            dyn.robot.handle, th0, dth0, tm, _host.gravity(g), fm, float(dt), int(intRes), self._jl)
        outs = [pos, vel, acc]
        if single:
            outs = [o[0] for o in outs]
        if not on_dev:
            outs = [_host.to_host(o) for o in outs]
        self._tick(t0, transfers=0 if on_dev else 7, kernel="forward_dynamics_rollout")
        return {"positions": outs[0], "velocities": outs[1], "accelerations": outs[2]}


    # -- small planner utilities kept for drop-in use -------------------------------------------------------
    def calculate_derivatives(self, positions, dt):
        """First, second and third finite differences of a sampled trajectory
        (planning/trajectory_dynamics.py:710-731): ``(v, a, j)`` with one row less each.  A few
        subtractions per row: done where the data lives (NumPy on the host, torch on the device)."""
        if _host.is_device_tensor(positions):
            v = (positions[1:] - positions[:-1]) / dt
            a = (v[1:] - v[:-1]) / dt
            return v, a, (a[1:] - a[:-1]) / dt
        x = np.asarray(positions)
        v = (x[1:] - x[:-1]) / dt
        a = (v[1:] - v[:-1]) / dt
        return v, a, (a[1:] - a[:-1]) / dt

    def cleanup_gpu_memory(self) -> None:
        """The reference frees its per-instance device arrays and memory pool
        (planning/trajectory_planning.py:502-530); here device buffers live only for the duration
        of a call, so this just returns torch's cached blocks to the driver."""
        torch.cuda.empty_cache()


def _input_is_f32(x) -> bool:
    if isinstance(x, torch.Tensor):
        return x.dtype == torch.float32
    return np.asarray(x).dtype == np.float32


# reference alias (planning/trajectory_planning.py:804-817)
TrajectoryPlanning = OptimizedTrajectoryPlanning
