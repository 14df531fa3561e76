"""``SerialManipulator`` -- drop-in mirror of ``ManipulaPy.kinematics.SerialManipulator``
(kinematics/serial_manipulator.py:43-162) for the batched space-frame hot path.

``forward_kinematics`` (kinematics/fk.py:39-86) and ``jacobian``
(kinematics/jacobian.py:39-93) accept a single ``(n,)`` configuration (reference behaviour:
returns ``(4, 4)`` / ``(6, n)`` float64) or a batch ``(P, n)`` (returns ``(P, 4, 4)`` /
``(P, 6, n)``), in the space or the body frame.  Both run in hand-written CUDA kernels; there
is no CPU path.  ``iterative_inverse_kinematics`` (kinematics/ik.py:39-311, all modes) runs
one target per thread in a batched damped-least-squares kernel; ``smart_`` / ``robust_inverse_kinematics``
(:327-598) are host restart logic around it (``ik_helpers``); ``trac_ik`` is out of scope.

Not a blanket drop-in for the reference classes: ``OptimizedTrajectoryPlanning`` needs THIS
package's ``ManipulatorDynamics`` (a reference object is wrapped automatically through
``ManipulatorDynamics.from_reference``) and raises on ``use_cuda=False``.
"""

from __future__ import annotations

from typing import Any, Optional

import numpy as np
import torch

from . import _host, _native


def _adjoint(T: np.ndarray) -> np.ndarray:
    """6x6 adjoint [[R, 0], [[p] R, R]] for twists [w; v] (utils/se3.py:45-52)."""
    R, p = T[:3, :3], T[:3, 3]
    px = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
    A = np.zeros((6, 6))
    A[:3, :3] = R
    A[3:, 3:] = R
    A[3:, :3] = px @ R
    return A


def _screws_from_axes(omega_list, r_list) -> np.ndarray:
    """S = [w; -w x r] per joint (reference utils/screw.py extract_screw_list semantics)."""
    w = np.asarray(omega_list, dtype=np.float64)
    r = np.asarray(r_list, dtype=np.float64)
    if w.ndim == 1:
        w = w.reshape(3, -1, order="F")
    if r.ndim == 1:
        r = r.reshape(3, -1, order="F")
    v = -np.cross(w.T, r.T).T
    return np.vstack([w, v])


class RobotHandle:
    """Owns one ``mpk_robot`` (the constant pack in joint-aligned frames, host resident)."""

    def __init__(self, S_list, M, Glist=None, Mlist_per_link=None, flags: int = 0):
        S = np.ascontiguousarray(np.asarray(S_list, dtype=np.float64))
        if S.ndim != 2 or S.shape[0] != 6:
            raise ValueError(f"S_list must be (6, n), got {S.shape}")
        Mh = np.asarray(M, dtype=np.float64)
        if Mh.ndim == 3:  # stack of poses: the reference uses the last one (kinematics/fk.py:69)
            Mh = Mh[-1]
        G = None if Glist is None else np.ascontiguousarray(np.asarray(Glist, dtype=np.float64))
        Mc = None if Mlist_per_link is None else np.ascontiguousarray(
            np.asarray(Mlist_per_link, dtype=np.float64))
        ops = _native.ops()
        self.n = int(S.shape[1])
        self._ops = ops
        self.handle = ops.robot_create(
            torch.from_numpy(S), torch.from_numpy(np.ascontiguousarray(Mh)),
            None if G is None else torch.from_numpy(G), None if Mc is None else torch.from_numpy(Mc), flags)
        self.rigid = bool(ops.robot_is_rigid(self.handle))

    def __del__(self):
        h, self.handle = getattr(self, "handle", 0), 0
        if h:
            try:
                self._ops.robot_destroy(h)
            except Exception:
                pass


def restart_noise_table(n: int, max_iterations: int):
    """Standard normals for the stagnation restarts of ONE iterative_inverse_kinematics run, drawn
    from NumPy's global generator exactly as the reference would draw them (``np.random.randn(n)``
    per restart, kinematics/ik.py:208), for as many restarts as the iteration budget allows (one
    per 21 iterations).  Returns ``(table (rows, n), generator state before the draws)``."""
    state = np.random.get_state()
    rows = max(1, int(max_iterations) // 21 + 1)
    return np.stack([np.random.randn(n) for _ in range(rows)]), state


def settle_generator(state, n: int, restarts: int) -> None:
    """Leave NumPy's global generator where the reference would: advanced by ``restarts`` draws."""
    np.random.set_state(state)
    for _ in range(int(restarts)):
        np.random.randn(n)


class SerialManipulator:
    """Kinematic model of a serial manipulator (space-frame product of exponentials)."""

    def __init__(self, M_list, omega_list=None, r_list=None, b_list=None, S_list=None, B_list=None,
                 G_list=None, joint_limits=None, *, device: Optional[Any] = None):
        self.M_list = np.asarray(M_list, dtype=np.float64)
        self.G_list = G_list
        self.omega_list = omega_list
        if S_list is None:
            if omega_list is None or r_list is None:
                raise ValueError("either S_list or (omega_list, r_list) is required")
            S_list = _screws_from_axes(omega_list, r_list)
        self.S_list = np.asarray(S_list, dtype=np.float64)
        self.B_list = None if B_list is None else np.asarray(B_list, dtype=np.float64)
        self.r_list = r_list
        self.b_list = b_list
        n = self.S_list.shape[1]
        self.joint_limits = joint_limits if joint_limits is not None else [(None, None)] * n
        self._device_arg = device
        self._robot: Optional[RobotHandle] = None
        self._robot_body: Optional[RobotHandle] = None

    # -- native handle ---------------------------------------------------------------------
    def _make_robot(self) -> RobotHandle:
        return RobotHandle(self.S_list, self.M_list)

    @property
    def robot(self) -> RobotHandle:
        if self._robot is None:
            self._robot = self._make_robot()
        return self._robot

    def _home_pose(self) -> np.ndarray:
        M = np.asarray(self.M_list, dtype=np.float64)
        return M[-1] if M.ndim == 3 else M

    def _body_screws(self) -> np.ndarray:
        """``B_list`` as the reference derives it (kinematics/serial_manipulator.py:75-95): given,
        or from (omega_list, b_list), or -- with neither -- the screws consistent with S_list,
        ``B_i = Ad(M^-1) S_i``."""
        if self.B_list is not None:
            return self.B_list
        if self.b_list is not None and self.omega_list is not None:
            return _screws_from_axes(self.omega_list, self.b_list)
        return _adjoint(np.linalg.inv(self._home_pose())) @ self.S_list

    @property
    def robot_body(self) -> RobotHandle:
        """Kinematics-only handle of the chain with space screws ``S'_i = Ad(M) B_i``:
        ``M prod e^{[B_i] th_i} = prod e^{[S'_i] th_i} M`` and ``J_b = Ad(T^-1) J_s'``."""
        if self._robot_body is None:
            M = self._home_pose()
            self._robot_body = RobotHandle(_adjoint(M) @ self._body_screws(), M)
        return self._robot_body

    @property
    def device(self) -> torch.device:
        return _host.default_device(self._device_arg)

    @property
    def num_joints(self) -> int:
        return int(self.S_list.shape[1])

    def _rows(self, x, name: str, keep_f32: bool = True):
        """-> (device tensor (P, n), single?, on_device?)"""
        on_dev = _host.is_device_tensor(x)
        t = _host.to_device(x, x.device if on_dev else self.device, keep_f32=keep_f32)
        single = t.dim() == 1
        n = self.num_joints
        if t.shape[-1] != n:
            raise ValueError(f"{name} must have {n} joint values per row, got shape {tuple(t.shape)}")
        return t.reshape(-1, n), single, on_dev

    @staticmethod
    def _finish(t: torch.Tensor, single: bool, on_dev: bool):
        if single:
            t = t[0]
        return t if on_dev else _host.to_host(t)

    # -- hot path ---------------------------------------------------------------------------
    def forward_kinematics(self, thetalist, frame: str = "space", precision=None):
        """End-effector pose(s) ``T = prod_i exp([S_i] theta_i) M`` (kinematics/fk.py:61-70).

        ``precision="float32"`` (extension) runs the float32 kernel and returns float32 arrays."""
        if frame not in ("space", "body"):
            raise ValueError("Invalid frame specified. Choose 'space' or 'body'.")
        th, single, on_dev = self._rows(thetalist, "thetalist")
        rb = self.robot if frame == "space" else self.robot_body  # same pose law, screws Ad(M) B
        T, _ = _native.ops().fk_jacobian(rb.handle, th, True, False, _host.is_f32(precision), False)
        return self._finish(T, single, on_dev)

    def jacobian(self, thetalist, frame: str = "space", precision=None):
        """Space Jacobian(s) ``J[:, i] = Ad(prod_{j<i} exp([S_j] theta_j)) S_i`` (kinematics/jacobian.py:62-73)."""
        if frame not in ("space", "body"):
            raise ValueError("Invalid frame specified. Choose 'space' or 'body'.")
        th, single, on_dev = self._rows(thetalist, "thetalist")
        body = frame == "body"
        rb = self.robot_body if body else self.robot
        _, J = _native.ops().fk_jacobian(rb.handle, th, False, True, _host.is_f32(precision), body)
        return self._finish(J, single, on_dev)

    def end_effector_velocity(self, thetalist, dthetalist, frame: str = "space"):
        """End-effector twist ``J(theta) dtheta`` in the space or body frame (kinematics/velocity.py:39-63);
        ``(n,)`` inputs -> ``(6,)``, batched ``(P, n)`` -> ``(P, 6)``.  The Jacobians come from the
        FK / Jacobian kernel and stay on the device for the product."""
        if frame not in ("space", "body"):
            raise ValueError("Invalid frame specified. Choose 'space' or 'body'.")
        th, single, on_dev = self._rows(thetalist, "thetalist")
        dth, _, _ = self._rows(dthetalist, "dthetalist")
        if dth.shape != th.shape:
            raise ValueError("thetalist and dthetalist must have the same shape")
        body = frame == "body"
        rb = self.robot_body if body else self.robot
        _, J = _native.ops().fk_jacobian(rb.handle, th, False, True, False, body)
        V = torch.einsum("prn,pn->pr", J, dth.to(J.dtype))
        return self._finish(V, single, on_dev or _host.is_device_tensor(dthetalist))

    def iterative_inverse_kinematics(self, T_desired, thetalist0, eomg: float = 1e-6, ev: float = 1e-6,
                                     max_iterations: int = 10000, plot_residuals: bool = False,
                                     damping: float = 2e-2, step_cap: float = 0.3,
                                     png_name: str = "ik_residuals.png", weight_orientation: float = 1.0,
                                     weight_position: float = 1.0, adaptive_tuning: bool = False,
                                     backtracking: bool = False, *, seed: int = 0):
        """Damped-least-squares IK with step cap, joint-limit projection, best-iterate tracking
        and stagnation restart (kinematics/ik.py:39-311), optionally with the reference's
        Levenberg-Marquardt adaptation (``adaptive_tuning``) and line search (``backtracking``).

        Reference call: ``T_desired (4, 4)``, ``thetalist0 (n,)`` -> ``(theta (n,), success, iterations)``.
        Batched extension: ``(P, 4, 4)``, ``(P, n)`` -> ``(theta (P, n), success (P,) bool,
        iterations (P,) int32)``, one target per GPU thread.  ``plot_residuals`` is not part of the
        kernel.  Stagnation restarts (a 0.1 sigma kick off the best iterate after 20 iterations
        without progress): a single-target call on host arrays takes that noise from NumPy's global
        generator draw for draw like the reference, and leaves the generator where the reference
        would; batched and device-resident calls use a counter-based generator keyed by ``seed``."""
        if plot_residuals:
            raise NotImplementedError("plot_residuals is outside the B200 hot path (no plotting)")
        if int(max_iterations) < 1:
            # the reference's loop never runs and its epilogue touches the unset loop counter (kinematics/ik.py:271-273)
            raise UnboundLocalError("max_iterations must be at least 1 (the reference fails on its unset loop counter 'k')")
        flags = (1 if adaptive_tuning else 0) | (2 if backtracking else 0)
        on_dev = _host.any_device(T_desired, thetalist0)
        dev = (thetalist0.device if _host.is_device_tensor(thetalist0)
               else T_desired.device if _host.is_device_tensor(T_desired) else self.device)
        th0 = _host.to_device(thetalist0, dev)
        single = th0.dim() == 1
        n = self.num_joints
        th0 = th0.reshape(-1, n)
        Td = _host.to_device(T_desired, dev).reshape(-1, 4, 4)
        if Td.shape[0] != th0.shape[0]:
            raise ValueError(f"{Td.shape[0]} target poses for {th0.shape[0]} initial guesses")
        lim = np.empty((n, 2))
        for i in range(n):
            mn, mx = (self.joint_limits[i] if i < len(self.joint_limits) else (None, None))
            lim[i] = (-np.inf if mn is None else mn, np.inf if mx is None else mx)
        noise = state = None
        if single and not on_dev and int(max_iterations) > 20:
            table, state = restart_noise_table(n, int(max_iterations))
            noise = torch.from_numpy(table).to(dev).reshape(1, -1, n)
        theta, ok, it, restarts = _native.ops().inverse_kinematics_dls(
            self.robot.handle, Td, th0, float(eomg), float(ev), int(max_iterations), float(damping), float(step_cap),
            float(weight_orientation), float(weight_position), torch.from_numpy(lim), int(seed), True, flags, noise)
        if state is not None:
            settle_generator(state, n, int(restarts[0].item()))
        ok = ok.bool()
        if single:
            th1 = theta[0] if on_dev else _host.to_host(theta[0])
            return th1, bool(ok[0].item()), int(it[0].item())
        if on_dev:
            return theta, ok, it
        return _host.to_host(theta), ok.cpu().numpy(), it.cpu().numpy()

    # -- IK front ends (kinematics/ik.py:327-598): initial-guess strategies + restarts around the kernel --
    def _ik_limits(self):
        n = self.num_joints
        lim = list(self.joint_limits)[:n]
        return lim + [(None, None)] * (n - len(lim))

    def _ik_batch(self, T_desired):
        if _host.is_device_tensor(T_desired):
            T_desired = T_desired.detach().cpu().numpy()
        Td = np.asarray(T_desired, dtype=np.float64)
        return Td.reshape(-1, 4, 4), Td.ndim == 2

    def smart_inverse_kinematics(self, T_desired, strategy: str = "workspace_heuristic", theta_current=None,
                                 T_current=None, cache=None, eomg: float = 1e-6, ev: float = 1e-6,
                                 max_iterations: int = 10000, plot_residuals: bool = False, damping: float = 2e-2,
                                 step_cap: float = 0.3, png_name: str = "ik_residuals.png",
                                 weight_orientation: float = 1.0, weight_position: float = 1.0,
                                 adaptive_tuning: bool = True, backtracking: bool = True,
                                 auto_fallback: bool = True, *, seed: int = 0):
        """Initial guess by ``strategy`` (workspace_heuristic, midpoint, random), then
        ``iterative_inverse_kinematics``; with ``auto_fallback`` up to four more starts (midpoint,
        3 x random) for targets that failed, keeping the best iterate (kinematics/ik.py:327-475).
        ``T_desired (4, 4)`` -> ``(theta, success, iterations)`` like the reference; batched
        extension ``(P, 4, 4)`` -> arrays, each fall-back round one launch over the failed targets.
        The 'extrapolate' and 'cached' strategies are host-side bookkeeping outside the hot path."""
        from . import ik_helpers

        valid = ["workspace_heuristic", "extrapolate", "cached", "random", "midpoint"]
        if strategy not in valid:
            raise ValueError(f"Unknown strategy '{strategy}'. Choose from: {valid}")
        if strategy == "extrapolate" and theta_current is not None and T_current is not None:
            raise NotImplementedError("strategy 'extrapolate' is outside the B200 hot path")
        if strategy == "cached" and cache is not None:
            raise NotImplementedError("strategy 'cached' is outside the B200 hot path")
        if strategy in ("extrapolate", "cached"):
            strategy = "workspace_heuristic"  # the reference's fall-back when the inputs are missing (:431-434)
        Td, single = self._ik_batch(T_desired)

        def solve(Tds, th0):
            args = (eomg, ev, max_iterations, plot_residuals, damping, step_cap, png_name, weight_orientation,
                    weight_position, adaptive_tuning, backtracking)
            if single:  # the reference's call: restart noise from NumPy's generator, like the guesses
                th, ok, it = self.iterative_inverse_kinematics(Tds[0], th0[0], *args)
                return th[None], np.array([ok]), np.array([it])
            return self.iterative_inverse_kinematics(Tds, th0, *args, seed=seed)

        theta, ok, it = ik_helpers.smart_driver(solve, self.forward_kinematics, Td, self.num_joints,
                                                self._ik_limits(), strategy, auto_fallback)
        if single:
            return theta[0], bool(ok[0]), int(it[0])
        return theta, ok, it

    def robust_inverse_kinematics(self, T_desired, max_attempts: int = 10, eomg: float = 2e-3, ev: float = 2e-3,
                                  max_iterations: int = 5000, verbose: bool = False, *, seed: int = 0):
        """Multi-start IK (kinematics/ik.py:477-598): up to ``max_attempts`` of ten (guess, damping,
        step cap) combinations with adaptive tuning and the line search on, tracking the best
        iterate.  Returns ``(theta, success, total_iterations, winning_strategy)``; batched
        extension ``(P, 4, 4)`` -> arrays (``winning_strategy`` an object array of names)."""
        from . import ik_helpers

        Td, single = self._ik_batch(T_desired)

        def solve(Tds, th0, damping, step_cap):
            kw = dict(damping=damping, step_cap=step_cap, adaptive_tuning=True, backtracking=True)
            if single:
                th, ok, it = self.iterative_inverse_kinematics(Tds[0], th0[0], eomg, ev, max_iterations, **kw)
                return th[None], np.array([ok]), np.array([it])
            return self.iterative_inverse_kinematics(Tds, th0, eomg, ev, max_iterations, seed=seed, **kw)

        theta, ok, it, win = ik_helpers.robust_driver(solve, self.forward_kinematics, Td, self.num_joints,
                                                      self._ik_limits(), max_attempts)
        if verbose:
            print(f"robust_inverse_kinematics: {int(ok.sum())} of {ok.size} targets solved")
        if single:
            return theta[0], bool(ok[0]), int(it[0]), str(win[0])
        return theta, ok, it, win

    def forward_kinematics_and_jacobian(self, thetalist, precision=None, frame: str = "space"):
        """Both outputs from one fused kernel launch (batched extension)."""
        if frame not in ("space", "body"):
            raise ValueError("Invalid frame specified. Choose 'space' or 'body'.")
        th, single, on_dev = self._rows(thetalist, "thetalist")
        body = frame == "body"
        rb = self.robot_body if body else self.robot
        T, J = _native.ops().fk_jacobian(rb.handle, th, True, True, _host.is_f32(precision), body)
        return self._finish(T, single, on_dev), self._finish(J, single, on_dev)
