"""Initial-guess strategies and the multi-start drivers of the reference's IK front ends, batched.

Mirrors ``ManipulaPy.kinematics.ik_helpers`` (workspace_heuristic_guess :28-113, random_in_limits
:179-212, midpoint_of_limits :215-246) and the restart logic of
``SerialManipulator.smart_inverse_kinematics`` / ``robust_inverse_kinematics``
(kinematics/ik.py:327-598).  The guesses are a handful of host flops per target; the solves and
the pose checks run on the GPU through the callables the drivers are given, and only the targets
that have not converged yet are solved again in each fall-back round (one batched launch each).

Random guesses come from NumPy's global generator, drawn joint by joint like the reference
(``np.random.uniform``); for a batch they are drawn target by target in index order, so a
single-target call consumes the generator exactly like the reference does.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

Limits = Sequence[Tuple[Optional[float], Optional[float]]]


def _clip_to_limits(theta: np.ndarray, joint_limits: Limits) -> np.ndarray:
    """(ik_helpers.py:407-446) rows of ``theta (..., n)`` clipped to the limits that are set."""
    n = theta.shape[-1]
    lower, upper = np.full(n, -np.inf), np.full(n, np.inf)
    for i, (mn, mx) in enumerate(list(joint_limits)[:n]):
        if mn is not None:
            lower[i] = mn
        if mx is not None:
            upper[i] = mx
    return np.minimum(np.maximum(theta, lower), upper).astype(theta.dtype)


def workspace_heuristic_guess(T_desired, n_joints: int, joint_limits: Limits) -> np.ndarray:
    """Geometric guess (ik_helpers.py:28-113): the first joints from the target position, the wrist
    from its rotation.  ``T_desired (4, 4)`` -> ``(n,)``; ``(P, 4, 4)`` -> ``(P, n)``."""
    T = np.asarray(T_desired, dtype=np.float64)
    single = T.ndim == 2
    T = T.reshape(-1, 4, 4)
    P = T.shape[0]
    ang = np.zeros((P, n_joints))
    p, R = T[:, :3, 3], T[:, :3, :3]
    if n_joints >= 1:
        ang[:, 0] = np.arctan2(p[:, 1], p[:, 0])
    if n_joints >= 2:
        r_xy = np.sqrt(p[:, 0] ** 2 + p[:, 1] ** 2)
        ang[:, 1] = np.where(r_xy > 1e-6, np.arctan2(p[:, 2], r_xy), 0.0)
    if n_joints >= 3:
        ang[:, 2] = np.pi / 4
    if n_joints > 3:
        gen = np.abs(R[:, 2, 2]) < 0.9999
        if n_joints >= 4:
            ang[:, 3] = np.where(gen, np.arctan2(R[:, 1, 2], R[:, 0, 2]), np.arctan2(R[:, 1, 0], R[:, 0, 0]))
        if n_joints >= 5:
            ang[:, 4] = np.where(gen, np.arccos(np.clip(R[:, 2, 2], -1, 1)), 0.0)
        if n_joints >= 6:
            ang[:, 5] = np.where(gen, np.arctan2(R[:, 2, 1], -R[:, 2, 0]), 0.0)
    ang = _clip_to_limits(ang, joint_limits)
    return ang[0] if single else ang


def random_in_limits(joint_limits: Limits) -> np.ndarray:
    """One uniform configuration within the limits (ik_helpers.py:179-212), NumPy's global generator."""
    angles = []
    for mn, mx in joint_limits:
        if mn is not None and mx is not None:
            angles.append(np.random.uniform(mn, mx))
        elif mn is not None:
            angles.append(mn + np.random.uniform(0, np.pi))
        elif mx is not None:
            angles.append(mx - np.random.uniform(0, np.pi))
        else:
            angles.append(np.random.uniform(-np.pi, np.pi))
    return np.asarray(angles, dtype=np.float64)


def random_in_limits_batch(joint_limits: Limits, count: int) -> np.ndarray:
    """``count`` configurations, the same values and the same use of NumPy's global generator as
    ``count`` calls of random_in_limits (one vectorised draw when every joint has both limits)."""
    lim = list(joint_limits)
    if count and all(mn is not None and mx is not None for mn, mx in lim):
        lo = np.array([mn for mn, _ in lim], dtype=np.float64)
        hi = np.array([mx for _, mx in lim], dtype=np.float64)
        return np.random.uniform(lo, hi, (count, len(lim)))
    return np.stack([random_in_limits(lim) for _ in range(count)]) if count else np.empty((0, len(lim)))


def midpoint_of_limits(joint_limits: Limits) -> np.ndarray:
    """(ik_helpers.py:215-246)"""
    return np.asarray([(mn + mx) / 2.0 if mn is not None and mx is not None else 0.0 for mn, mx in joint_limits],
                      dtype=np.float64)


def pose_error(T_curr: np.ndarray, T_desired: np.ndarray) -> np.ndarray:
    """Position + rotation-angle distance of poses ``(P, 4, 4)`` (kinematics/ik.py:313-325)."""
    pos = np.linalg.norm(T_curr[:, :3, 3] - T_desired[:, :3, 3], axis=1)
    tr = np.einsum("pji,pji->p", T_curr[:, :3, :3], T_desired[:, :3, :3])  # trace(Rc^T Rd)
    return pos + np.arccos(np.clip((tr - 1) / 2, -1, 1))


def _guess(strategy: str, Td: np.ndarray, n: int, limits: Limits) -> np.ndarray:
    """Initial guesses ``(len(Td), n)`` of one strategy."""
    if strategy == "workspace_heuristic":
        return workspace_heuristic_guess(Td, n, limits).reshape(-1, n)
    if strategy == "midpoint":
        return np.tile(midpoint_of_limits(limits)[:n], (Td.shape[0], 1))
    if strategy == "random":
        return random_in_limits_batch(limits, Td.shape[0])[:, :n]
    raise NotImplementedError(
        f"initial-guess strategy '{strategy}' is not part of the batched front end "
        "(workspace_heuristic, midpoint and random are)")


Solve = Callable[..., Tuple[np.ndarray, np.ndarray, np.ndarray]]
ForwardKinematics = Callable[[np.ndarray], np.ndarray]


def smart_driver(solve: Solve, fk: ForwardKinematics, Td: np.ndarray, n: int, limits: Limits, strategy: str,
                 auto_fallback: bool) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Restart logic of smart_inverse_kinematics (ik.py:436-475) over a batch ``Td (P, 4, 4)``:
    the primary strategy for every target, then up to four fall-back rounds (midpoint, 3 x random)
    for the targets that failed, keeping the iterate with the smallest pose error.
    ``solve(Td_subset, theta0_subset) -> (theta, success, iterations)``; ``fk(theta) -> poses``."""
    theta, ok, iters = solve(Td, _guess(strategy, Td, n, limits))
    theta, ok, total = np.array(theta, dtype=np.float64), np.array(ok, dtype=bool), np.array(iters, dtype=np.int64)
    if not auto_fallback or ok.all():
        return theta, ok, total
    best = pose_error(fk(theta), Td)
    for fallback in ("midpoint", "random", "random", "random"):
        idx = np.flatnonzero(~ok)
        if idx.size == 0:
            break
        th_try, ok_try, it_try = solve(Td[idx], _guess(fallback, Td[idx], n, limits))
        ok_try = np.asarray(ok_try, dtype=bool)
        total[idx] += np.asarray(it_try, dtype=np.int64)
        err = pose_error(fk(th_try), Td[idx])
        take = ok_try | (err < best[idx])
        theta[idx[take]] = np.asarray(th_try)[take]
        best[idx[take]] = err[take]
        ok[idx[ok_try]] = True
    return theta, ok, total


ROBUST_STRATEGIES: List[Tuple[str, float, float]] = [  # (guess, damping, step_cap), ik.py:505-516
    ("workspace_heuristic", 0.02, 0.3), ("midpoint", 0.02, 0.3), ("workspace_heuristic", 0.01, 0.4),
    ("random", 0.02, 0.3), ("random", 0.03, 0.25), ("midpoint", 0.01, 0.4), ("random", 0.015, 0.35),
    ("random", 0.025, 0.3), ("workspace_heuristic", 0.03, 0.25), ("random", 0.02, 0.35),
]


def robust_driver(solve: Solve, fk: ForwardKinematics, Td: np.ndarray, n: int, limits: Limits,
                  max_attempts: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Multi-start of robust_inverse_kinematics (ik.py:518-598) over a batch: attempt k uses guess,
    damping and step cap number k of ROBUST_STRATEGIES on the targets that are still unsolved.
    ``solve(Td_subset, theta0_subset, damping=..., step_cap=...)``.  Returns theta, success, total
    iterations and the name of the winning (or best failing) strategy per target."""
    P = Td.shape[0]
    theta = np.tile(midpoint_of_limits(limits)[:n], (P, 1))
    ok = np.zeros(P, dtype=bool)
    total = np.zeros(P, dtype=np.int64)
    best = np.full(P, np.inf)
    winner = np.array(["none"] * P, dtype=object)
    for name, damping, step_cap in ROBUST_STRATEGIES[: max(0, min(max_attempts, len(ROBUST_STRATEGIES)))]:
        idx = np.flatnonzero(~ok)
        if idx.size == 0:
            break
        th_try, ok_try, it_try = solve(Td[idx], _guess(name, Td[idx], n, limits), damping=damping, step_cap=step_cap)
        ok_try = np.asarray(ok_try, dtype=bool)
        total[idx] += np.asarray(it_try, dtype=np.int64)
        err = pose_error(fk(th_try), Td[idx])
        take = ok_try | (err < best[idx])
        theta[idx[take]] = np.asarray(th_try)[take]
        best[idx[take]] = err[take]
        winner[idx[take]] = name
        ok[idx[ok_try]] = True
    return theta, ok, total, winner
