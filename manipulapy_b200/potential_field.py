"""``CollisionChecker`` / ``PotentialField`` -- mirrors of ``ManipulaPy.potential_field`` for the
collision / limit post-processing hook of ``joint_trajectory``.

The reference runs, for EVERY trajectory row on the host: ``URDF.link_fk`` over the link tree,
the axis-aligned-box overlap test of every pair of link hulls outside the allowed-collision set
(potential_field/collision.py:162-221) and, for colliding rows, up to 100 potential-field steps
``row -= 0.01 * gradient`` (planning/collision_host.py:40-88).  Here one CUDA thread owns one row
(``csrc/collision.cu``): a whole trajectory -- or a batch of them -- is checked and nudged in one
launch.

The checker is built from DATA, not from a URDF (URDF / mesh parsing is outside the hot path):

``links``   the link table the reference's URDF loader extracts (``robots/<name>_links.npz``, written
            by ``oracle/gen_collision_golden.py``): ``link_names``, ``link_joint`` (the actuated joint a
            link hangs on, -1 = fixed to the base), ``link_home`` (poses at the zero configuration),
            ``link_acm`` (allowed-collision matrix: parent / child and grandparent pairs);
``hulls``   ``{link name: (V, 3) points in the link frame}`` -- what ``ConvexHull(...).points`` holds
            in the reference (its hulls come from meshes, which this package does not load).
"""

from __future__ import annotations

from typing import Any, Dict, Iterable, Mapping, Optional

import numpy as np
import torch

from . import _host, _native


class PotentialField:
    """Gains of the artificial potential field (potential_field/fields.py:40-58).  The trajectory
    hook calls ``compute_gradient(row, thetaend, [])``: with no obstacles only the attractive term
    ``attractive_gain * (q - q_goal)`` is left, which is what the kernel applies; the repulsive term
    (:126-166) is provided here on the host for API parity."""

    def __init__(self, attractive_gain: float = 1.0, repulsive_gain: float = 100.0,
                 influence_distance: float = 0.5) -> None:
        self.attractive_gain = attractive_gain
        self.repulsive_gain = repulsive_gain
        self.influence_distance = influence_distance

    def compute_gradient(self, q, q_goal, obstacles: Iterable[Any]):
        q = np.asarray(q)
        diff = (q - np.asarray(q_goal)) * 1.0
        grad = np.asarray(self.attractive_gain, dtype=diff.dtype) * diff
        rep = np.zeros(q.shape, dtype=diff.dtype)
        for ob in obstacles:
            d_vec = (q - np.asarray(ob)) * 1.0
            d = np.linalg.norm(d_vec)
            if d > self.influence_distance:
                continue
            if d < 1e-10:
                e = np.zeros(q.shape, dtype=diff.dtype)
                e[0] = 1.0
                rep = rep + self.repulsive_gain * e
                continue
            d_safe, d0 = max(d, 1e-10), max(self.influence_distance, 1e-10)
            rep = rep + (-40.0 * self.repulsive_gain * (1.0 / d_safe - 1.0 / d0) * (1.0 / d_safe ** 3)) * d_vec
        return grad + rep


class CollisionChecker:
    """Self-collision checker over link hulls (potential_field/collision.py:29-221), batched.

    ``dynamics``: this package's ``SerialManipulator`` / ``ManipulatorDynamics`` of the robot (its
    constant pack gives the joint frames the link poses hang on)."""

    def __init__(self, dynamics, links: Mapping[str, Any], hulls: Mapping[str, Any], *, device: Optional[Any] = None):
        self.dynamics = dynamics
        self.link_names = [str(x) for x in links["link_names"]]
        self.link_joint = np.ascontiguousarray(links["link_joint"], dtype=np.int32)
        self.link_home = np.ascontiguousarray(links["link_home"], dtype=np.float64)
        self.link_acm = np.ascontiguousarray(links["link_acm"], dtype=np.uint8)
        L = len(self.link_names)
        if self.link_home.shape != (L, 4, 4) or self.link_acm.shape != (L, L) or self.link_joint.shape != (L,):
            raise ValueError("links: link_joint (L,), link_home (L, 4, 4), link_acm (L, L) are required")
        # insertion order of the hull dict = the reference's iteration order (collision.py:180)
        self.convex_hulls: Dict[str, np.ndarray] = {}
        for name, pts in hulls.items():
            if name not in self.link_names:
                raise KeyError(f"hull for unknown link '{name}'")
            p = np.ascontiguousarray(getattr(pts, "points", pts), dtype=np.float64)
            if p.ndim != 2 or p.shape[1] != 3 or p.shape[0] < 1:
                raise ValueError(f"hull '{name}' must be (V, 3) points")
            self.convex_hulls[name] = p
        self.device = _host.default_device(device if device is not None else getattr(dynamics, "_device_arg", None))
        ops = _native.ops()
        names = list(self.convex_hulls)
        hull_link = torch.tensor([self.link_names.index(nm) for nm in names], dtype=torch.int32)
        hull_count = torch.tensor([self.convex_hulls[nm].shape[0] for nm in names], dtype=torch.int32)
        pts = (torch.from_numpy(np.concatenate([self.convex_hulls[nm] for nm in names])) if names
               else torch.zeros((0, 3), dtype=torch.float64))
        self._model_host = ops.collision_model_pack(
            dynamics.robot.handle, torch.from_numpy(self.link_joint), torch.from_numpy(self.link_home),
            torch.from_numpy(self.link_acm), hull_link, hull_count, pts)
        self._model_dev = self._model_host.to(self.device)

    # -- helpers ----------------------------------------------------------------------------------
    def _rows(self, thetalist):
        on_dev = _host.is_device_tensor(thetalist)
        t = _host.to_device(thetalist, thetalist.device if on_dev else self.device, keep_f32=True)
        single = t.dim() == 1
        n = self.dynamics.num_joints
        if t.shape[-1] != n:
            raise ValueError(f"configurations must have {n} joint values, got shape {tuple(t.shape)}")
        return t.reshape(-1, n), single, on_dev

    # -- API --------------------------------------------------------------------------------------
    def check_collision(self, thetalist):
        """``(n,)`` configuration -> ``bool`` (the reference call); ``(P, n)`` -> uint8 flags ``(P,)``."""
        th, single, on_dev = self._rows(thetalist)
        if not self.convex_hulls:
            flags = torch.zeros(th.shape[0], dtype=torch.uint8, device=th.device)
        else:
            flags = _native.ops().self_collision(self.dynamics.robot.handle, self._model_host, self._model_dev, th)
        if single:
            return bool(flags[0].item())
        return flags if on_dev else flags.cpu().numpy()

    def link_fk_batch(self, cfgs, links=None):
        """``URDF.link_fk_batch`` (urdf/core.py:577-633): ``{link name: (P, 4, 4)}`` float64."""
        th, _, on_dev = self._rows(cfgs)
        T = _native.ops().link_fk_batch(self.dynamics.robot.handle, self._model_host, self._model_dev, th,
                                        len(self.link_names))
        out = {}
        for l, name in enumerate(self.link_names):
            if links is None or name in links:
                out[name] = T[:, l] if on_dev else T[:, l].cpu().numpy()
        return out

    def link_fk(self, cfg=None, links=None, use_names: bool = True):
        """``URDF.link_fk`` for one configuration: ``{link name: (4, 4)}``."""
        n = self.dynamics.num_joints
        cfg = np.zeros(n) if cfg is None else cfg
        return {k: v[0] for k, v in self.link_fk_batch(np.asarray(cfg, dtype=np.float64).reshape(1, n), links).items()}

    def avoid(self, rows, goal, potential_field: Optional[PotentialField] = None, *, rows_per_goal: Optional[int] = None,
              step: float = 0.01, max_iterations: int = 100, return_info: bool = False):
        """``_apply_collision_avoidance_cpu`` (planning/collision_host.py:40-88): float32 rows ``(P, n)``
        nudged towards ``goal`` (``(n,)``, or ``(B, n)`` with ``rows_per_goal`` rows each) while they
        collide.  Device tensors are modified in place; host arrays are copied."""
        pf = potential_field or PotentialField()
        on_dev = _host.is_device_tensor(rows)
        n = self.dynamics.num_joints
        if on_dev:
            r = rows.reshape(-1, n)
            if r.dtype != torch.float32 or not r.is_contiguous():
                raise ValueError("device rows must be contiguous float32")
        else:
            r = torch.from_numpy(np.ascontiguousarray(np.asarray(rows, dtype=np.float32).reshape(-1, n))).to(self.device)
        g = torch.as_tensor(np.asarray(goal.cpu() if isinstance(goal, torch.Tensor) else goal, dtype=np.float32)).reshape(-1, n)
        P = r.shape[0]
        rpg = int(rows_per_goal) if rows_per_goal is not None else max(1, -(-P // g.shape[0]))
        if not self.convex_hulls or P == 0:
            it = torch.zeros(P, dtype=torch.int32, device=r.device)
            fl = torch.zeros(P, dtype=torch.uint8, device=r.device)
        else:
            it, fl = _native.ops().collision_avoidance(self.dynamics.robot.handle, self._model_host, self._model_dev, r,
                                                       g.to(r.device), rpg, float(pf.attractive_gain), float(step),
                                                       int(max_iterations))
        out = r.reshape(rows.shape) if on_dev else r.cpu().numpy().reshape(np.asarray(rows).shape)
        if return_info:
            return out, (it if on_dev else it.cpu().numpy()), (fl if on_dev else fl.cpu().numpy())
        return out
