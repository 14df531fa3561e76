"""Operator registry -- the drop-in seam of ``ManipulaPy.cuda_kernels``
(cuda_kernels/registry.py:46-89, 828-952).

Same shape as the reference: a frozen ``KernelRegistration`` per name inside a
``KernelRegistry`` whose ``register`` refuses duplicates (``ValueError``), whose ``get``
raises ``KeyError("Unknown CUDA kernel '<name>'. Available kernels: ...")`` and whose
``execute`` runs the launcher.  The seven ``trajectory.*`` names of the reference all map
onto the ONE hand-written trajectory kernel (the variants differed only in launch shape);
the dynamics / kinematics kernels, which the reference never registered, are added under
``dynamics.*`` / ``kinematics.*``.  There is no CPU launcher: ``cpu_launcher`` raises.
"""

from __future__ import annotations

from dataclasses import dataclass
from types import MappingProxyType
from typing import Any, Callable, Dict, Mapping

import numpy as np

from . import _host, _native


@dataclass(frozen=True)
class KernelRegistration:
    name: str
    implementation: Any
    launch_config: Callable[..., Any]
    cpu_fallback: Callable[..., Any]
    gpu_launcher: Callable[..., Any]
    cpu_launcher: Callable[..., Any]
    metadata: Mapping[str, Any]

    def __post_init__(self) -> None:
        object.__setattr__(self, "metadata", MappingProxyType(dict(self.metadata)))


class KernelRegistry:
    def __init__(self) -> None:
        self._entries: Dict[str, KernelRegistration] = {}

    def register(self, entry: KernelRegistration) -> None:
        if entry.name in self._entries:
            raise ValueError(f"CUDA kernel '{entry.name}' is already registered")
        self._entries[entry.name] = entry

    def get(self, name: str) -> KernelRegistration:
        try:
            return self._entries[name]
        except KeyError:
            available = ", ".join(sorted(self._entries))
            raise KeyError(f"Unknown CUDA kernel '{name}'. Available kernels: {available}") from None

    def names(self):
        return sorted(self._entries)

    def execute(self, name: str, *args: Any, **kwargs: Any) -> Any:
        return self.get(name).gpu_launcher(*args, **kwargs)


def _no_cpu(*_a, **_k):
    raise RuntimeError("manipulapy_b200 has no CPU launcher; use the reference ManipulaPy on CPU")


def _launch_cfg_points(points: int, threads: int = 128):
    """(grid, block) of the one-thread-per-point kernels."""
    return ((int(points) + threads - 1) // threads,), (threads,)


def _launch_trajectory_gpu(thetastart, thetaend, Tf, N, method, use_pinned=True, *, enable_monitoring=True):
    """Reference launcher signature (registry.py:828-867): host ``(n,)`` endpoints ->
    ``(pos, vel, acc)`` host float32 ``(N, n)``; endpoints are rounded to float32 first.

    Registry contract (cuda_kernels/trajectory_kernels.py:40-76, 179, 195-198): any method other
    than 3 / 5 is LINEAR, and ``N <= 1`` or ``Tf <= 0`` sits at the start configuration -- unlike
    the planner's CPU kernel (zero scaling / NaN), which ``joint_trajectory`` mirrors."""
    dev = _host.default_device()
    if int(N) <= 0:
        n = np.asarray(thetastart).shape[-1]
        return tuple(np.zeros((0, n), np.float32) for _ in range(3))
    s = _host.to_device(np.asarray(thetastart, dtype=np.float32), dev).reshape(1, -1)
    e = _host.to_device(np.asarray(thetaend, dtype=np.float32), dev).reshape(1, -1)
    pos, vel, acc = _native.ops().joint_trajectory(s, e, True, float(Tf), int(N),
                                                   (int(method) & 0xFF) | REGISTRY_CONTRACT, None)
    return tuple(_host.to_host(x[0]) for x in (pos, vel, acc))


REGISTRY_CONTRACT = 0x100  # MPK_TRAJ_REGISTRY_CONTRACT (include/mpk.h)

KERNEL_REGISTRY = KernelRegistry()

for _variant in ("auto", "auto_tune", "standard", "vectorized", "memory_optimized", "warp_optimized",
                 "cache_friendly"):
    KERNEL_REGISTRY.register(KernelRegistration(
        name=f"trajectory.{_variant}", implementation="mpk_joint_trajectory",
        launch_config=_launch_cfg_points, cpu_fallback=_no_cpu, gpu_launcher=_launch_trajectory_gpu,
        cpu_launcher=_no_cpu,
        metadata={"c_abi": "mpk_joint_trajectory", "source": "csrc/traj.cu", "variant_of": "trajectory"}))

for _name, _sym, _src in (
    ("kinematics.fk_jacobian_space", "mpk_fk_jacobian_space", "csrc/kin.cu"),
    ("dynamics.inverse", "mpk_inverse_dynamics", "csrc/dyn.cu"),
    ("dynamics.trajectory_inverse", "mpk_trajectory_inverse_dynamics", "csrc/dyn.cu"),
    ("dynamics.mass_matrix", "mpk_mass_matrix", "csrc/dyn.cu"),
    ("dynamics.forward", "mpk_forward_dynamics", "csrc/dyn.cu"),
    ("dynamics.forward_rollout", "mpk_forward_dynamics_trajectory", "csrc/fd_flavour.cu"),
    ("kinematics.inverse_dls", "mpk_inverse_kinematics_dls", "csrc/ik.cu"),
    ("trajectory.cartesian", "mpk_cartesian_trajectory", "csrc/traj.cu"),
):
    def _make(sym):
        def _launch(*args, **kwargs):
            op = {"mpk_fk_jacobian_space": "fk_jacobian", "mpk_inverse_dynamics": "inverse_dynamics",
                  "mpk_trajectory_inverse_dynamics": "trajectory_inverse_dynamics",
                  "mpk_mass_matrix": "mass_matrix", "mpk_forward_dynamics": "forward_dynamics",
                  "mpk_forward_dynamics_trajectory": "forward_dynamics_trajectory",
                  "mpk_inverse_kinematics_dls": "inverse_kinematics_dls",
                  "mpk_cartesian_trajectory": "cartesian_trajectory"}[sym]
            return getattr(_native.ops(), op)(*args, **kwargs)
        return _launch

    KERNEL_REGISTRY.register(KernelRegistration(
        name=_name, implementation=_sym, launch_config=_launch_cfg_points, cpu_fallback=_no_cpu,
        gpu_launcher=_make(_sym), cpu_launcher=_no_cpu, metadata={"c_abi": _sym, "source": _src}))


def execute_registered_kernel(name: str, *args: Any, **kwargs: Any) -> Any:
    """Reference entry point (cuda_kernels/registry.py:963-965)."""
    return KERNEL_REGISTRY.execute(name, *args, **kwargs)


def check_cuda_availability() -> bool:
    import torch

    return torch.cuda.is_available()
