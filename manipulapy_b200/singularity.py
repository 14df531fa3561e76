"""``Singularity`` -- batched mirror of the Jacobian / forward-kinematics callers in
``ManipulaPy.singularity`` (singularity/singularity_analysis.py:40-330).

The reference evaluates these one configuration at a time on the host; here the space Jacobians
(or the end-effector positions) of a whole batch come from the FK / Jacobian kernel and the
6 x n singular values from ``torch.linalg.svdvals`` on the device (a library call on a 6 x n
matrix per configuration -- not part of the hot path).  Plotting is out of scope:
``plot_workspace_monte_carlo`` (:165-243) becomes ``workspace_points``, which returns the sampled
end-effector positions (and their convex hull on request) instead of drawing them.
"""

from __future__ import annotations

from typing import Any, List, Sequence, Tuple

import numpy as np
import torch

from . import _host


class Singularity:
    def __init__(self, serial_manipulator: Any) -> None:
        self.serial_manipulator = serial_manipulator

    def _singular_values(self, thetalist):
        """-> (singular values (P, min(6, n)) descending on the device, single?, on_device?)"""
        sm = self.serial_manipulator
        on_dev = _host.is_device_tensor(thetalist)
        J = sm.jacobian(thetalist if on_dev else _host.to_device(thetalist, sm.device), frame="space")
        single = J.dim() == 2
        return torch.linalg.svdvals(J.reshape(-1, 6, J.shape[-1])), single, on_dev

    def singularity_analysis(self, thetalist):
        """True where the smallest singular value of the space Jacobian is below 1e-4 (:52-74);
        a ``(P, n)`` batch gives a bool array."""
        sv, single, on_dev = self._singular_values(thetalist)
        flag = sv[:, -1] < 1e-4
        if single:
            return bool(flag[0].item())
        return flag if on_dev else flag.cpu().numpy()

    def condition_number(self, thetalist):
        """sigma_max / sigma_min of the space Jacobian, inf for a rank-deficient one (:246-286)."""
        sv, single, on_dev = self._singular_values(thetalist)
        ratio = sv[:, 0] / sv[:, -1]
        ratio = torch.where(torch.isnan(ratio), torch.full_like(ratio, float("inf")), ratio)
        if single:
            return float(ratio[0].item())
        return ratio if on_dev else ratio.cpu().numpy()

    def near_singularity_detection(self, thetalist, threshold: float = 1e-2):
        """condition_number > threshold (:288-305; the reference's default threshold is kept)."""
        c = self.condition_number(thetalist)
        return c > threshold

    def workspace_points(self, joint_limits: Sequence[Tuple[float, float]], num_samples: int = 10000, *,
                         seed: int = 1234, return_samples: bool = False, return_hull: bool = False):
        """Monte-Carlo workspace estimate (:165-243 without the plot): ``num_samples`` joint vectors
        uniform within ``joint_limits`` (float32 like the reference's sampler, drawn on the device by
        torch's Philox generator -- the reference uses Numba's CUDA xoroshiro128+, so the sample
        set differs), their end-effector positions ``(num_samples, 3)`` from ONE batched forward
        kinematics launch, optionally the samples and the ``scipy.spatial.ConvexHull`` of the points."""
        sm = self.serial_manipulator
        dev = sm.device
        lim = torch.as_tensor(np.asarray(joint_limits, dtype=np.float32), device=dev)
        gen = torch.Generator(device=dev).manual_seed(int(seed))
        u = torch.rand(int(num_samples), lim.shape[0], dtype=torch.float32, device=dev, generator=gen)
        samples = u * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
        T = sm.forward_kinematics(samples)  # float32 rows are upcast exactly by the kernel
        pts = _host.to_host(T[:, :3, 3].contiguous())
        out: List[Any] = [pts]
        if return_samples:
            out.append(_host.to_host(samples))
        if return_hull:
            from scipy.spatial import ConvexHull

            out.append(ConvexHull(pts))
        return out[0] if len(out) == 1 else tuple(out)
