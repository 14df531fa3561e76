"""``ManipulatorDynamics`` -- drop-in mirror of ``ManipulaPy.dynamics.ManipulatorDynamics``
(dynamics/manipulator_dynamics.py:43-86) for the batched dynamics hot path.

Every method takes one ``(n,)`` sample (reference behaviour: float64 result of the
reference's shape) or a batch ``(P, n)``:

=============================  ===============================  ==========================
method                         reference                        result
=============================  ===============================  ==========================
``mass_matrix``                dynamics/mass_matrix.py:16-99    ``(n, n)`` / ``(P, n, n)``
``velocity_quadratic_forces``  dynamics/forces.py:26-59         ``(n,)`` / ``(P, n)``
``gravity_forces``             dynamics/forces.py:61-133        ``(n,)`` / ``(P, n)``
``inverse_dynamics``           dynamics/id_fd.py:16-48          ``(n,)`` / ``(P, n)``
``forward_dynamics``           dynamics/id_fd.py:50-83          ``(n,)`` / ``(P, n)``
=============================  ===============================  ==========================

The reference evaluates ``tau = M ddth + c + g + Js^T Ftip`` with a finite-difference
Coriolis term (dynamics/cache.py:23-56, eps = 1e-6); the kernels evaluate the same
quantity analytically (Newton-Euler recursion over the link-CoM inertias), so results agree
with the reference to its own finite-difference noise (~1e-9 absolute) and with the
reference's mass matrix / gravity to rounding (1e-15).

The LEGACY path of the reference (``Mlist_per_link is None``: dynamics/mass_matrix.py:101-132,
dynamics/forces.py:135-154; documented there as incorrect physics) is available behind an explicit
opt-in, ``ManipulatorDynamics(..., Mlist_per_link=None, legacy=True)``: the same five methods then
return the reference's legacy numbers (``csrc/legacy.cu``, literal formulas incl. the
finite-difference Coriolis term).  Without the flag a missing ``Mlist_per_link`` raises.
"""

from __future__ import annotations

from typing import Any, Optional

import numpy as np

from . import _host, _native
from .kinematics import RobotHandle, SerialManipulator


class ManipulatorDynamics(SerialManipulator):
    def __init__(self, M_list, omega_list=None, r_list=None, b_list=None, S_list=None, B_list=None,
                 Glist=None, Mlist_per_link=None, *, device: Optional[Any] = None,
                 force_general_inertia: bool = False, legacy: bool = False):
        super().__init__(M_list, omega_list, r_list, b_list, S_list, B_list, device=device)
        if Glist is None:
            raise ValueError("Glist is required")
        self._legacy = False
        if Mlist_per_link is None:
            if not legacy:
                raise NotImplementedError(
                    "Mlist_per_link=None selects the reference's legacy dynamics path (dynamics/mass_matrix.py:"
                    "101-132, documented there as incorrect).  Pass Mlist_per_link for the per-link model, or "
                    "opt in with legacy=True to get the reference's legacy numbers")
            import warnings

            warnings.warn("ManipulatorDynamics without Mlist_per_link: using the reference's legacy approximation "
                          "(incorrect for non-trivial robots)", stacklevel=2)
            self._legacy = True
        self.Glist = np.asarray(Glist, dtype=np.float64)
        self.Mlist_per_link = None if Mlist_per_link is None else np.asarray(Mlist_per_link, dtype=np.float64)
        n = self.num_joints
        if self.Glist.shape != (n, 6, 6) or (self.Mlist_per_link is not None and self.Mlist_per_link.shape != (n, 4, 4)):
            raise ValueError("Glist must be (n, 6, 6) and Mlist_per_link (n, 4, 4)")
        self._force_general = bool(force_general_inertia)

    @classmethod
    def from_reference(cls, dynamics, **kw) -> "ManipulatorDynamics":
        """Wrap a reference ``ManipulatorDynamics`` (or anything exposing its constant pack)."""
        return cls(dynamics.M_list, getattr(dynamics, "omega_list", None), getattr(dynamics, "r_list", None),
                   getattr(dynamics, "b_list", None), dynamics.S_list, getattr(dynamics, "B_list", None),
                   dynamics.Glist, dynamics.Mlist_per_link, **kw)

    def _make_robot(self) -> RobotHandle:
        if self._legacy:
            return RobotHandle(self.S_list, self.M_list)  # kinematics only: the legacy dynamics take raw S, M, G
        return RobotHandle(self.S_list, self.M_list, self.Glist, self.Mlist_per_link,
                           flags=1 if self._force_general else 0)

    def _legacy_call(self, mode: int, th, dth=None, third=None, g=None, Ftip=None):
        """csrc/legacy.cu: mode 0 mass matrix, 1 gravity, 2 Coriolis, 3 inverse, 4 forward dynamics."""
        import torch

        P = th.shape[0]
        ftip, rows = self._ftip(Ftip, P, th.device)
        f64 = lambda t: None if t is None else t.to(torch.float64)  # noqa: E731
        return _native.ops().legacy_dynamics(
            torch.from_numpy(self.S_list), torch.from_numpy(self._home_pose()), torch.from_numpy(self.Glist), mode,
            f64(th), f64(dth), f64(third), [0.0, 0.0, 0.0] if g is None else g, ftip, rows)

    # -- helpers ------------------------------------------------------------------------------
    def _ftip(self, Ftip, P: int, device):
        """-> (single wrench list | None, per-point device rows | None)"""
        if Ftip is None:
            return None, None
        if _host.is_device_tensor(Ftip):
            if Ftip.dim() == 1:
                return _host.vec(Ftip, 6, "Ftip"), None
            return None, _host.to_device(Ftip, device).reshape(P, 6)
        a = np.asarray(Ftip, dtype=np.float64)
        if a.ndim == 1:
            return _host.vec(a, 6, "Ftip"), None
        if a.shape != (P, 6):
            raise ValueError(f"per-point Ftip must be ({P}, 6), got {a.shape}")
        return None, _host.to_device(a, device)

    def _id(self, th, dth, ddth, g, Ftip, out_f32=False, limits=None, precision=None):
        P = th.shape[0]
        ftip, rows = self._ftip(Ftip, P, th.device)
        return _native.ops().inverse_dynamics(self.robot.handle, th, dth, ddth, g, ftip, rows, limits, out_f32,
                                              _host.is_f32(precision))

    def computed_torque(self, thetalistd, dthetalistd, ddthetalistd, thetalist, dthetalist, g, Kp, Ki, Kd,
                        eint=None):
        """Torque of the reference's computed-torque law (control/computed_torque.py:60-95), batched, for
        callers that keep the integral state ``eint`` themselves:
        ``tau = M(theta) (Kp e + Ki eint + Kd de) + inverse_dynamics(theta, dtheta, ddtheta_d, g, 0)``.
        Inverse dynamics is affine in the acceleration with slope ``M(theta)``, so this is ONE inverse
        dynamics call at the acceleration ``ddtheta_d + Kp e + Ki eint + Kd de`` -- no mass matrix is formed."""
        th, single, on_dev = self._rows(thetalist, "thetalist", keep_f32=False)
        rows = lambda x: self._rows(x, "row", keep_f32=False)[0]  # noqa: E731
        dth, thd, dthd, ddthd = rows(dthetalist), rows(thetalistd), rows(dthetalistd), rows(ddthetalistd)
        gain = lambda k: _host.to_device(np.broadcast_to(np.asarray(k, dtype=np.float64), (self.num_joints,)).copy(),  # noqa: E731
                                         th.device)
        e = thd - th
        a = ddthd + gain(Kp) * e + gain(Kd) * (dthd - dth)
        if eint is not None:
            a = a + gain(Ki) * rows(eint)
        if self._legacy:
            return self._finish(self._legacy_call(3, th, dth, a, _host.gravity(g), None), single, on_dev)
        return self._finish(self._id(th, dth, a, _host.gravity(g), None), single, on_dev)

    # -- hot path ---------------------------------------------------------------------------------
    def mass_matrix(self, thetalist):
        th, single, on_dev = self._rows(thetalist, "thetalist")
        M = self._legacy_call(0, th) if self._legacy else _native.ops().mass_matrix(self.robot.handle, th)
        return self._finish(M, single, on_dev)

    def velocity_quadratic_forces(self, thetalist, dthetalist, precision=None):
        th, single, on_dev = self._rows(thetalist, "thetalist")
        dth, _, _ = self._rows(dthetalist, "dthetalist")
        th, dth = _host.promote_rows(th, dth)
        if self._legacy:
            return self._finish(self._legacy_call(2, th, dth), single, on_dev)
        c = self._id(th, dth, None, [0.0, 0.0, 0.0], None, precision=precision)
        return self._finish(c, single, on_dev)

    def gravity_forces(self, thetalist, g=None, precision=None):
        th, single, on_dev = self._rows(thetalist, "thetalist")
        if self._legacy:
            return self._finish(self._legacy_call(1, th, g=_host.gravity(g)), single, on_dev)
        out = self._id(th, None, None, _host.gravity(g), None, precision=precision)
        return self._finish(out, single, on_dev)

    def inverse_dynamics(self, thetalist, dthetalist, ddthetalist, g=None, Ftip=None, precision=None):
        """``precision="float32"`` (extension): float32 arithmetic, 1e-4 relative on torques."""
        th, single, on_dev = self._rows(thetalist, "thetalist")
        dth, _, _ = self._rows(dthetalist, "dthetalist")
        ddth, _, _ = self._rows(ddthetalist, "ddthetalist")
        th, dth, ddth = _host.promote_rows(th, dth, ddth)
        if self._legacy:
            return self._finish(self._legacy_call(3, th, dth, ddth, _host.gravity(g), Ftip), single, on_dev)
        tau = self._id(th, dth, ddth, _host.gravity(g), Ftip, precision=precision)
        return self._finish(tau, single, on_dev)

    def forward_dynamics(self, thetalist, dthetalist, taulist, g=None, Ftip=None):
        th, single, on_dev = self._rows(thetalist, "thetalist", keep_f32=False)
        dth, _, _ = self._rows(dthetalist, "dthetalist", keep_f32=False)
        tau, _, _ = self._rows(taulist, "taulist", keep_f32=False)
        if self._legacy:
            return self._finish(self._legacy_call(4, th, dth, tau, _host.gravity(g), Ftip), single, on_dev)
        ftip, rows = self._ftip(Ftip, th.shape[0], th.device)
        dd = _native.ops().forward_dynamics(self.robot.handle, th, dth, tau, _host.gravity(g), ftip, rows)
        return self._finish(dd, single, on_dev)
