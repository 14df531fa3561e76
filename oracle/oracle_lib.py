"""ctypes binding of ``oracle.c`` (the CPU restatement of the reference path).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.
"""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Optional

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)


class _Robot(C.Structure):
    _fields_ = [("n", C.c_int), ("S", _dp), ("M", _dp), ("G", _dp), ("Mcom", _dp)]


def build_oracle(force: bool = False) -> Path:
    """Compile ``oracle.c`` into ``oracle/_build/liboracle.so`` (gcc)."""
    src = _HERE / "oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s", "-B", "all"], check=True)
    return _LIB_PATH


_lib = None


def load_oracle():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build_oracle()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.orc_forward_dynamics.restype = C.c_int
        _lib.orc_forward_dynamics_analytic.restype = C.c_int
        _lib.orc_forward_dynamics_analytic_ldlt.restype = C.c_int
        _lib.orc_forward_dynamics_trajectory.restype = C.c_int
        _lib.orc_forward_dynamics_rollout_batch.restype = C.c_int
        _lib.orc_max_dof.restype = C.c_int
    return _lib


def _d(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _f(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a: Optional[np.ndarray], ty=_dp):
    return None if a is None else a.ctypes.data_as(ty)


_THREADS = 1


def set_threads(n: int) -> None:
    """Host threads used by the batched oracle loops (ctypes releases the GIL)."""
    global _THREADS
    _THREADS = max(1, int(n))


def _parallel(count: int, fn) -> None:
    """Run ``fn(lo, hi)`` over contiguous slices of ``range(count)`` on ``_THREADS`` threads."""
    if _THREADS == 1 or count < 2 * _THREADS:
        fn(0, count)
        return
    from concurrent.futures import ThreadPoolExecutor

    step = -(-count // _THREADS)
    with ThreadPoolExecutor(_THREADS) as ex:
        futs = [ex.submit(fn, lo, min(count, lo + step)) for lo in range(0, count, step)]
        for f in futs:
            f.result()


class Oracle:
    """The reference algorithm for one robot, float64, on the host.

    Parameters mirror the reference's constant pack
    (``dynamics/manipulator_dynamics.py:46-75``): ``S_list (6, n)``, ``M (4, 4)``,
    ``Glist (n, 6, 6)``, ``Mlist_per_link (n, 4, 4)``.
    """

    def __init__(self, S_list, M, Glist, Mlist_per_link):
        self.lib = load_oracle()
        self.S = _d(S_list)
        self.M = _d(M)
        self.G = _d(Glist)
        self.Mcom = _d(Mlist_per_link)
        self.n = int(self.S.shape[1])
        assert self.S.shape == (6, self.n) and self.M.shape == (4, 4)
        assert self.G.shape == (self.n, 6, 6) and self.Mcom.shape == (self.n, 4, 4)
        assert self.n <= self.lib.orc_max_dof()
        self._rb = _Robot(self.n, _p(self.S), _p(self.M), _p(self.G), _p(self.Mcom))
        self._r = C.byref(self._rb)

    # -- per-point float64 API (batched over the leading axis) -------------------
    def forward_kinematics(self, theta):
        th = _d(theta).reshape(-1, self.n)
        T = np.empty((th.shape[0], 4, 4))
        _parallel(th.shape[0], lambda lo, hi: self.lib.orc_fk_jacobian_batch(
            self._r, C.c_int64(hi - lo), _p(th[lo:hi]), _p(T[lo:hi]), None))
        return T

    def jacobian(self, theta):
        th = _d(theta).reshape(-1, self.n)
        J = np.empty((th.shape[0], 6, self.n))
        _parallel(th.shape[0], lambda lo, hi: self.lib.orc_fk_jacobian_batch(
            self._r, C.c_int64(hi - lo), _p(th[lo:hi]), None, _p(J[lo:hi])))
        return J

    def mass_matrix(self, theta, analytic=False):
        th = _d(theta).reshape(-1, self.n)
        M = np.empty((th.shape[0], self.n, self.n))
        _parallel(th.shape[0], lambda lo, hi: self.lib.orc_mass_matrix_batch(
            self._r, C.c_int64(hi - lo), _p(th[lo:hi]), C.c_int(int(analytic)), _p(M[lo:hi])))
        return M

    def inverse_dynamics(self, theta, dtheta, ddtheta, g=(0.0, 0.0, -9.81), Ftip=None, analytic=False):
        th = _d(theta).reshape(-1, self.n)
        P = th.shape[0]
        dth = _d(dtheta).reshape(P, self.n)
        ddth = _d(ddtheta).reshape(P, self.n)
        g = _d(g)
        if Ftip is None:
            F, stride = None, 0
        else:
            F = _d(Ftip)
            stride = 0 if F.ndim == 1 else 6
        out = np.empty((P, self.n))
        def run(lo, hi):
            Fs = None if F is None else (F if stride == 0 else F[lo:hi])
            self.lib.orc_inverse_dynamics_batch(
                self._r, C.c_int64(hi - lo), _p(th[lo:hi]), _p(dth[lo:hi]), _p(ddth[lo:hi]), _p(g),
                _p(Fs), C.c_int64(stride), C.c_int(int(analytic)), _p(out[lo:hi]))

        _parallel(P, run)
        return out

    def gravity_forces(self, theta, g=(0.0, 0.0, -9.81), analytic=False):
        th = _d(theta).reshape(-1, self.n)
        if analytic:
            z = np.zeros_like(th)
            return self.inverse_dynamics(th, z, z, g, None, analytic=True)
        out = np.empty_like(th)
        g = _d(g)
        for p in range(th.shape[0]):
            self.lib.orc_gravity_forces(self._r, _p(th[p]), _p(g), _p(out[p]))
        return out

    def velocity_quadratic_forces(self, theta, dtheta, analytic=False):
        th = _d(theta).reshape(-1, self.n)
        dth = _d(dtheta).reshape(-1, self.n)
        if analytic:
            return self.inverse_dynamics(th, dth, np.zeros_like(th), (0.0, 0.0, 0.0), None, analytic=True)
        out = np.empty_like(th)
        for p in range(th.shape[0]):
            self.lib.orc_velocity_quadratic_forces(self._r, _p(th[p]), _p(dth[p]), _p(out[p]))
        return out

    def forward_dynamics(self, theta, dtheta, tau, g=(0.0, 0.0, -9.81), Ftip=None, analytic=False):
        th = _d(theta).reshape(-1, self.n)
        P = th.shape[0]
        dth = _d(dtheta).reshape(P, self.n)
        ta = _d(tau).reshape(P, self.n)
        g = _d(g)
        F = np.zeros((P, 6)) if Ftip is None else np.broadcast_to(_d(Ftip), (P, 6)).copy()
        out = np.empty((P, self.n))
        fn = (self.lib.orc_forward_dynamics_analytic_ldlt if analytic == 2 else
              self.lib.orc_forward_dynamics_analytic if analytic else self.lib.orc_forward_dynamics)
        for p in range(P):
            rc = fn(self._r, _p(th[p]), _p(dth[p]), _p(ta[p]), _p(g), _p(F[p]), _p(out[p]))
            if rc:
                raise np.linalg.LinAlgError("singular mass matrix")
        return out

    # -- trajectory-level float32 API --------------------------------------------
    @staticmethod
    def joint_trajectory(start, end, Tf, N, method, joint_limits=None, inputs_f32=None):
        """``joint_trajectory`` (1-D start/end, float32-cast inputs) or
        ``batch_joint_trajectory`` ((B, n) start/end, input dtype preserved)."""
        lib = load_oracle()
        s = np.asarray(start)
        single = s.ndim == 1
        if inputs_f32 is None:
            inputs_f32 = single or s.dtype == np.float32
        s = _d(s).reshape(-1, s.shape[-1])
        e = _d(end).reshape(s.shape)
        B, n = s.shape
        lim = None if joint_limits is None else _f(joint_limits).reshape(n, 2)
        pos = np.empty((B, N, n), np.float32)
        vel = np.empty_like(pos)
        acc = np.empty_like(pos)
        for b in range(B):
            lib.orc_joint_trajectory(
                C.c_int(n), _p(s[b]), _p(e[b]), C.c_int(int(inputs_f32)), C.c_double(float(Tf)),
                C.c_int64(int(N)), C.c_int(int(method)), _p(lim, _fp), _p(pos[b], _fp),
                _p(vel[b], _fp), _p(acc[b], _fp))
        if single:
            pos, vel, acc = pos[0], vel[0], acc[0]
        return {"positions": pos, "velocities": vel, "accelerations": acc}

    # -- Cartesian straight-line trajectory (numpy restatement) ---------------------------------
    @staticmethod
    def _log3_vec(E):
        """Rotation vector of utils/so3.py:172-191 (MatrixLog3): atan2 angle (:150-169), Taylor-safe
        theta / sin theta (:114-147), half-turn axis from the symmetric part (:33-79)."""
        c = float(np.clip((np.trace(E) - 1) / 2, -1.0, 1.0))
        vee = np.array([E[2, 1] - E[1, 2], E[0, 2] - E[2, 0], E[1, 0] - E[0, 1]])
        theta = np.arctan2(np.sqrt(max(float(vee @ vee), 1e-300)) / 2, c)
        if theta > np.pi - 1e-2:
            sym = 0.5 * (E + E.T) - c * np.eye(3)
            j = 2 if sym[2, 2] >= 1e-6 else (1 if sym[1, 1] >= 1e-6 else 0)
            cand = sym[:, j]
            axis = cand / np.sqrt(max(float(cand @ cand), 1e-24))
            return theta * (1.0 if vee[j] >= 0 else -1.0) * axis
        if c > 1 - 5e-5:
            u = 1.0 - c
            coef = 1.0 + u / 3.0 + u * u * (4.0 / 45.0)
        else:
            cs = float(np.clip(c, -1.0 + 1e-7, 1.0 - 1e-7))
            coef = np.arccos(cs) / np.sqrt(max(1 - cs * cs, 1e-30))
        return 0.5 * coef * vee

    @staticmethod
    def cartesian_trajectory(Xstart, Xend, Tf, N, method):
        """``cartesian_trajectory`` restated from planning/trajectory.py:504-594 (orientation /
        position assembly) and :676-740 (linear velocity / acceleration), utils/so3.py:199-237
        (MatrixExp3) and utils/time_scaling.py:28-53.  float32 outputs like the reference."""
        Xs, Xe = _d(Xstart).reshape(4, 4), _d(Xend).reshape(4, 4)
        N = int(N)
        Rs, ps, pe = Xs[:3, :3], Xs[:3, 3], Xe[:3, 3]
        w = Oracle._log3_vec(Rs.T @ Xe[:3, :3])
        timegap = Tf / (N - 1.0)
        ori, pos, vel, acc = [], [], [], []
        for i in range(N):
            t = timegap * i
            s = 3 * (t / Tf) ** 2 - 2 * (t / Tf) ** 3 if method == 3 else \
                10 * (t / Tf) ** 3 - 15 * (t / Tf) ** 4 + 6 * (t / Tf) ** 5
            k = w * s
            th2 = float(k @ k)
            if th2 < 1e-4:
                A, B = 1.0 - th2 / 6.0 + th2 * th2 / 120.0, 0.5 - th2 / 24.0 + th2 * th2 / 720.0
            else:
                th = np.sqrt(max(th2, 1e-12))
                A, B = np.sin(th) / th, (1 - np.cos(th)) / (th * th)
            K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            ori.append(Rs @ (np.eye(3) + A * K + B * (K @ K)))
            pos.append(s * pe + (1 - s) * ps)
            tau = (i * (Tf / (N - 1))) / Tf
            if method == 3:
                sd, sdd = 6.0 * tau * (1.0 - tau) / Tf, 6.0 / (Tf * Tf) * (1.0 - 2.0 * tau)
            elif method == 5:
                t2, t3, t4 = tau * tau, tau * tau * tau, tau * tau * tau * tau
                sd, sdd = (30.0 * t2 - 60.0 * t3 + 30.0 * t4) / Tf, (60.0 * tau - 180.0 * t2 + 120.0 * t3) / (Tf * Tf)
            else:
                sd = sdd = 0.0
            vel.append(sd * (pe - ps))
            acc.append(sdd * (pe - ps))
        f = lambda a, tail: np.asarray(a, np.float32).reshape((N,) + tail)
        return {"positions": f(pos, (3,)), "velocities": f(vel, (3,)), "accelerations": f(acc, (3,)),
                "orientations": f(ori, (3, 3))}

    # -- inverse kinematics (numpy restatement; small cases only) -------------------------------
    def iterative_inverse_kinematics(self, T_desired, thetalist0, eomg=1e-6, ev=1e-6, max_iterations=10000,
                                     damping=2e-2, step_cap=0.3, weight_orientation=1.0, weight_position=1.0,
                                     joint_limits=None, adaptive_tuning=False, backtracking=False):
        """``iterative_inverse_kinematics`` restated from kinematics/ik.py:39-311: geometric error
        (:88-140), SVD damped least squares (:142-162), step cap and limit projection (:164-176,
        :253-262), best-iterate tracking and the stagnation restart drawn from NumPy's global
        generator (:196-213), the Levenberg-Marquardt damping / step-cap adaptation
        (``adaptive_tuning``, :215-229) and the five-scale line search (``backtracking``, :253-276).
        Returns (theta, success, iterations) like the reference."""
        theta = _d(thetalist0).copy()
        Td = _d(T_desired).reshape(4, 4)
        n = theta.shape[0]
        lo, hi = np.full(n, -np.inf), np.full(n, np.inf)
        if joint_limits is not None:
            for i, (mn, mx) in enumerate(list(joint_limits)[:n]):
                if mn is not None:
                    lo[i] = mn
                if mx is not None:
                    hi[i] = mx

        def err(T):
            pos = Td[:3, 3] - T[:3, 3]
            R = T[:3, :3]
            E = R.T @ Td[:3, :3]
            ang = np.arccos(np.clip((np.trace(E) - 1) / 2, -1, 1))
            vee = np.array([E[2, 1] - E[1, 2], E[0, 2] - E[2, 0], E[1, 0] - E[0, 1]])
            if ang < 1e-6:
                w = vee / 2
            elif abs(ang - np.pi) < 1e-6:
                w = ang * np.eye(3)[int(np.argmax(np.diag(E)))]
            else:
                w = ang * (vee / (2 * np.sin(ang) + 1e-10))
            return np.concatenate((R @ w, pos)), abs(ang), float(np.linalg.norm(pos))

        best_theta, best_error, stall, success = theta.copy(), np.inf, 0, False
        cur = np.inf
        damping_local, step_cap_local, prev_error, nu = damping, step_cap, np.inf, 2.0
        clip = lambda th: np.minimum(np.maximum(th, lo), hi)  # noqa: E731
        k = 0
        for k in range(max_iterations):
            V, rot, trans = err(self.forward_kinematics(theta)[0])
            cur = rot + trans
            if rot < eomg and trans < ev:
                success = True
                break
            if cur < best_error:
                best_error, best_theta, stall = cur, theta.copy(), 0
            else:
                stall += 1
            if stall > 20:
                theta = clip(best_theta + 0.1 * np.random.randn(n))
                damping_local, stall, nu = damping, 0, 2.0
                continue
            if adaptive_tuning and k > 0:
                if cur < prev_error * 0.75:
                    damping_local = max(1e-6, damping_local / 3)
                    step_cap_local = min(step_cap * 1.5, step_cap_local * 1.2)
                    nu = 2.0
                elif cur < prev_error * 0.95:
                    damping_local = max(1e-6, damping_local / 1.5)
                elif cur > prev_error:
                    damping_local = min(5e-1, damping_local * nu)
                    nu = min(nu * 1.5, 8)
                    step_cap_local = max(0.01, step_cap_local * 0.7)
            prev_error = cur
            J = self.jacobian(theta)[0]
            Vw = V * np.array([weight_orientation] * 3 + [weight_position] * 3)
            U, sv, Vt = np.linalg.svd(J, full_matrices=False)
            d = Vt.T @ ((sv / (sv**2 + damping_local**2 + 1e-12)) * (U.T @ Vw))
            nd = np.linalg.norm(d)
            if nd > step_cap_local:
                d = d * (step_cap_local / nd)
            if backtracking:
                best_scale_theta, best_scale_error = theta, cur
                for scale in (1.0, 0.5, 0.25, 0.125, 0.75):
                    cand = clip(theta + scale * d)
                    _, rot_t, trans_t = err(self.forward_kinematics(cand)[0])
                    if rot_t + trans_t < best_scale_error:
                        best_scale_error, best_scale_theta = rot_t + trans_t, cand
                theta = best_scale_theta if best_scale_error < cur * 1.1 else clip(theta + 0.1 * d)
            else:
                theta = clip(theta + d)
        else:
            k += 1 if max_iterations > 0 else 0
        if not success and best_error < cur:
            theta = best_theta
            _, rot, trans = err(self.forward_kinematics(theta)[0])
            success = bool(rot < eomg and trans < ev)
        return theta, success, k + 1

    # -- body frame (numpy restatement; small cases only) --------------------------------------
    @staticmethod
    def _exp_twist(S, th):
        """utils/se3.py:33-42 transform_from_twist (unit omega or omega = 0)."""
        w, v = S[:3], S[3:]
        T = np.eye(4)
        if np.linalg.norm(w) == 0.0:
            T[:3, 3] = v * th
            return T
        W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        R = np.eye(3) + np.sin(th) * W + (1 - np.cos(th)) * W @ W
        T[:3, :3] = R
        T[:3, 3] = (np.eye(3) * th + (1 - np.cos(th)) * W + (th - np.sin(th)) * W @ W) @ v
        return T

    @staticmethod
    def _adjoint(T):
        R, p = T[:3, :3], T[:3, 3]
        P = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
        A = np.zeros((6, 6))
        A[:3, :3] = R
        A[3:, 3:] = R
        A[3:, :3] = P @ R
        return A

    @staticmethod
    def body_forward_kinematics(M, B_list, theta):
        """T = M e^{[B_1] th_1} ... e^{[B_n] th_n}  (kinematics/fk.py:72-81)."""
        th = _d(theta).reshape(-1, B_list.shape[1])
        out = np.empty((th.shape[0], 4, 4))
        for p in range(th.shape[0]):
            T = np.eye(4)
            for i in range(B_list.shape[1]):
                T = T @ Oracle._exp_twist(B_list[:, i], th[p, i])
            out[p] = np.asarray(M, np.float64) @ T
        return out

    @staticmethod
    def body_jacobian(B_list, theta):
        """J_b[:, i] = Ad(e^{-[B_n] th_n} ... e^{-[B_{i+1}] th_{i+1}}) B_i  (kinematics/jacobian.py:74-90)."""
        n = B_list.shape[1]
        th = _d(theta).reshape(-1, n)
        out = np.empty((th.shape[0], 6, n))
        for p in range(th.shape[0]):
            T = np.eye(4)
            out[p, :, n - 1] = B_list[:, n - 1]
            for i in range(n - 2, -1, -1):
                T = T @ Oracle._exp_twist(B_list[:, i + 1], -th[p, i + 1])
                out[p, :, i] = Oracle._adjoint(T) @ B_list[:, i]
        return out

    @staticmethod
    def registry_trajectory(thetastart, thetaend, Tf, N, method):
        """The registry launchers' trajectory contract, restated from the reference's NumPy
        float32 fallback (cuda_kernels/trajectory_kernels.py:40-85): linear time scaling for a
        method other than 3 / 5, and s = ds = dds = 0 for N <= 1 or Tf <= 0.  float32 arithmetic
        throughout, like the reference (its CUDA kernels are float32 fast-math: parity with this
        path is "allclose", not bit-exact -- tests/test_cuda_kernels_cpu.py:227-262)."""
        s0 = np.asarray(thetastart, dtype=np.float32)
        e0 = np.asarray(thetaend, dtype=np.float32)
        N = int(N)
        if N <= 1 or Tf <= 0.0:
            s = np.zeros(N, np.float32)
            sd = np.zeros(N, np.float32)
            sdd = np.zeros(N, np.float32)
        else:
            t = np.linspace(0, Tf, N, dtype=np.float32)
            tau = t / Tf
            if method == 3:
                s = 3.0 * tau**2 - 2.0 * tau**3
                sd = 6.0 * tau * (1.0 - tau) / Tf
                sdd = 6.0 * (1.0 - 2.0 * tau) / (Tf * Tf)
            elif method == 5:
                s = 10.0 * tau**3 - 15.0 * tau**4 + 6.0 * tau**5
                sd = (30.0 * tau**2 - 60.0 * tau**3 + 30.0 * tau**4) / Tf
                sdd = (60.0 * tau - 180.0 * tau**2 + 120.0 * tau**3) / (Tf * Tf)
            else:
                s = tau
                sd = np.ones_like(tau) / Tf
                sdd = np.zeros_like(tau)
        d = e0 - s0
        pos = s0[None, :] + s[:, None] * d[None, :]
        return (pos.astype(np.float32), (sd[:, None] * d[None, :]).astype(np.float32),
                (sdd[:, None] * d[None, :]).astype(np.float32))

    def inverse_dynamics_trajectory(self, theta, dtheta, ddtheta, g=(0.0, 0.0, -9.81), Ftip=None,
                                    torque_limits=None, analytic=False):
        th = _d(theta).reshape(-1, self.n)
        P = th.shape[0]
        dth = _d(dtheta).reshape(P, self.n)
        ddth = _d(ddtheta).reshape(P, self.n)
        g = _d(g)
        F = np.zeros(6) if Ftip is None else _d(Ftip)
        lim = None if torque_limits is None else _f(torque_limits).reshape(self.n, 2)
        out = np.empty((P, self.n), np.float32)
        _parallel(P, lambda lo, hi: self.lib.orc_inverse_dynamics_trajectory(
            self._r, C.c_int64(hi - lo), _p(th[lo:hi]), _p(dth[lo:hi]), _p(ddth[lo:hi]), _p(g), _p(F),
            _p(lim, _fp), C.c_int(int(analytic)), _p(out[lo:hi], _fp)))
        return out

    def forward_dynamics_trajectory(self, theta0, dtheta0, taumat, g, Ftipmat, dt, intRes,
                                    joint_limits=None, analytic=False):
        """Single ``(n,)`` start or batched ``(B, n)`` starts with ``taumat (B, N, n)``.
        ``analytic``: False literal reference algorithm, True analytic recursion + LU, 2 analytic
        recursion + the kernels' LDL^T solve."""
        th0 = _d(theta0)
        single = th0.ndim == 1
        th0 = th0.reshape(-1, self.n)
        B = th0.shape[0]
        dth0 = _d(dtheta0).reshape(B, self.n)
        tm = _d(taumat).reshape(B, -1, self.n)
        N = tm.shape[1]
        if N == 0:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")
        g = _d(g)
        Fm = None if Ftipmat is None else _d(Ftipmat).reshape(B, N, 6)
        lim = None if joint_limits is None else _f(joint_limits).reshape(self.n, 2)
        pos = np.empty((B, N, self.n), np.float32)
        vel = np.empty_like(pos)
        acc = np.empty_like(pos)
        rcs = []

        def run(lo, hi):
            rcs.append(self.lib.orc_forward_dynamics_rollout_batch(
                self._r, C.c_int64(hi - lo), _p(th0[lo:hi]), _p(dth0[lo:hi]), C.c_int64(N),
                _p(tm[lo:hi]), _p(g), None if Fm is None else _p(Fm[lo:hi]), C.c_double(float(dt)),
                C.c_int(int(intRes)), _p(lim, _fp), C.c_int(int(analytic)), _p(pos[lo:hi], _fp),
                _p(vel[lo:hi], _fp), _p(acc[lo:hi], _fp)))

        _parallel(B, run)
        if any(rcs):
            raise np.linalg.LinAlgError("singular mass matrix during rollout")
        if single:
            pos, vel, acc = pos[0], vel[0], acc[0]
        return {"positions": pos, "velocities": vel, "accelerations": acc}


class CollisionOracle:
    """numpy restatement of the collision / limit post-processing hook of ``joint_trajectory``
    (SURVEY.md 8f-1).  TEST INFRASTRUCTURE ONLY; pinned on tests/golden/collision.npz, which records
    what the unmodified reference does with injected hulls (oracle/gen_collision_golden.py).

    ``link_joint[l]``   index of the actuated joint link l hangs on (-1: fixed to the base)
    ``link_home[l]``    pose of link l at the zero configuration
    ``hulls``           {link index: (V, 3) points in the link frame}, in the checker's insertion order
    ``acm[a, b]``       1 if the pair is excluded (adjacent / grandparent links,
                        potential_field/adjacency.py:9-30)
    """

    def __init__(self, S_list, link_joint, link_home, hulls, acm):
        self.S = _d(S_list)
        self.n = self.S.shape[1]
        self.link_joint = np.asarray(link_joint, np.int64)
        self.link_home = _d(link_home)
        self.hulls = {int(k): _d(v) for k, v in hulls.items()}
        self.acm = np.asarray(acm, bool)

    def link_fk_batch(self, cfgs):
        """URDF.link_fk_batch (urdf/core.py:577-633): pose of every link for every configuration.  The
        tree walk ``T_child = T_parent @ origin @ R(axis, q)`` equals, for the serial chain whose space
        screws the same loader extracted, ``prod_{j <= k} e^{[S_j] q_j} @ T_link(0)``."""
        q = _d(cfgs).reshape(-1, self.n)
        L = self.link_home.shape[0]
        out = np.empty((q.shape[0], L, 4, 4))
        for p in range(q.shape[0]):
            P = [np.eye(4)]
            for j in range(self.n):
                P.append(P[-1] @ Oracle._exp_twist(self.S[:, j], q[p, j]))
            for l in range(L):
                out[p, l] = P[self.link_joint[l] + 1] @ self.link_home[l]
        return out

    def check_collision(self, cfgs):
        """CollisionChecker.check_collision (potential_field/collision.py:162-195) per configuration:
        any pair of hulls, not in the allowed-collision set, whose world axis-aligned boxes overlap
        (``_points_intersect`` :197-221: max_a >= min_b and max_b >= min_a on every axis)."""
        q = np.asarray(cfgs)
        q = q.reshape(-1, self.n)
        T = self.link_fk_batch(q.astype(np.float64))
        names = list(self.hulls)
        flags = np.zeros(q.shape[0], np.uint8)
        for p in range(q.shape[0]):
            box = {}
            for l in names:
                w = (T[p, l, :3, :3] @ self.hulls[l].T + T[p, l, :3, 3:4]).T
                box[l] = (w.min(0), w.max(0))
            hit = False
            for ia in range(len(names)):
                for ib in range(ia + 1, len(names)):
                    a, b = names[ia], names[ib]
                    if self.acm[a, b]:
                        continue
                    if np.all(box[a][1] >= box[b][0]) and np.all(box[b][1] >= box[a][0]):
                        hit = True
                        break
                if hit:
                    break
            flags[p] = hit
        return flags

    def avoid(self, rows, goal, attractive_gain=1.0, step=0.01, max_iterations=100):
        """_apply_collision_avoidance_cpu (planning/collision_host.py:40-88) on float32 rows with the
        float32 goal ``thetaend``: a colliding row takes up to ``max_iterations`` steps
        ``row <- row - 0.01 * gradient`` in float32 arithmetic, where with no obstacles
        ``PotentialField.compute_gradient`` (potential_field/fields.py:112-170) is
        ``attractive_gain * (row - goal)``, until ``check_collision(row)`` clears.
        Returns (rows, iterations taken per row)."""
        out = np.array(rows, np.float32, copy=True)
        goal = np.asarray(goal, np.float32)
        gain, st = np.float32(attractive_gain), np.float32(step)
        iters = np.zeros(out.shape[0], np.int32)
        for i in range(out.shape[0]):
            r = out[i]
            if self.check_collision(r)[0]:
                for k in range(max_iterations):
                    grad = gain * ((r - goal) * np.float32(1.0))
                    r = np.asarray(r - st * grad, np.float32)
                    iters[i] = k + 1
                    if not self.check_collision(r)[0]:
                        break
            out[i] = r
        return out, iters


class LegacyOracle:
    """numpy restatement of the reference's LEGACY dynamics path (no ``Mlist_per_link``; SURVEY.md
    8f-4).  TEST INFRASTRUCTURE ONLY; pinned on tests/golden/legacy_dynamics.npz.

    With ``T_i = e^{[S_1] th_1} ... e^{[S_i] th_i} M`` (kinematics/fk.py:61-70 on ``theta[:i + 1]`` -- the
    END-EFFECTOR home pose for every link) and ``J`` the space Jacobian:
      mass matrix   row i = J[:, i]^T (Ad(T_i)^T G_i Ad(T_i)) J, then 0.5 (M + M^T)   (mass_matrix.py:101-132)
      gravity       g_i = (R_i^T g) . colsum(G_i[:3, :3])                             (forces.py:135-154)
      Coriolis      c_i = dth^T Gamma_i dth, Gamma from central differences of M, eps = 1e-6 (forces.py:26-59,
                    cache.py:23-56)
      inverse dynamics  M ddth + c + g + J^T Ftip;   forward dynamics  solve(M, tau - c - g - J^T Ftip)
    """

    def __init__(self, S_list, M, Glist):
        self.S, self.M, self.G = _d(S_list), _d(M), _d(Glist)
        self.n = self.S.shape[1]

    def _prefixes(self, th):
        T, out, J = np.eye(4), [], np.empty((6, self.n))
        for i in range(self.n):
            J[:, i] = Oracle._adjoint(T) @ self.S[:, i]
            T = T @ Oracle._exp_twist(self.S[:, i], th[i])
            out.append(T @ self.M)
        return out, J

    def mass_matrix(self, th):
        Ts, J = self._prefixes(_d(th))
        Mm = np.empty((self.n, self.n))
        for i in range(self.n):
            Ad = Oracle._adjoint(Ts[i])
            Mm[i] = (J[:, i] @ (Ad.T @ self.G[i] @ Ad)) @ J
        return 0.5 * (Mm + Mm.T)

    def gravity_forces(self, th, g):
        Ts, _ = self._prefixes(_d(th))
        return np.array([(Ts[i][:3, :3].T @ _d(g)) @ self.G[i][:3, :3].sum(0) for i in range(self.n)])

    def velocity_quadratic_forces(self, th, dth, eps=1e-6):
        th, dth, n = _d(th), _d(dth), self.n
        dM = np.empty((n, n, n))
        for k in range(n):
            e = np.zeros(n)
            e[k] = 1.0
            dM[:, :, k] = (self.mass_matrix(th + eps * e) - self.mass_matrix(th - eps * e)) / (2.0 * eps)
        return np.array([dth @ (0.5 * (dM[i] + dM[i].T - dM[:, :, i])) @ dth for i in range(n)])

    def inverse_dynamics(self, th, dth, ddth, g, Ftip):
        _, J = self._prefixes(_d(th))
        return (self.mass_matrix(th) @ _d(ddth) + self.velocity_quadratic_forces(th, dth) + self.gravity_forces(th, g)
                + J.T @ _d(Ftip))

    def forward_dynamics(self, th, dth, tau, g, Ftip):
        _, J = self._prefixes(_d(th))
        rhs = _d(tau) - self.velocity_quadratic_forces(th, dth) - self.gravity_forces(th, g) - J.T @ _d(Ftip)
        return np.linalg.solve(self.mass_matrix(th), rhs)
