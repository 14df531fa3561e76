"""Bundled robot constant packs.

Each ``<name>.npz`` holds exactly what the reference's ``URDFToSerialManipulator`` produces
for its bundled URDF (urdf/core.py:670-769): ``S_list (6, n)``, ``M (4, 4)``,
``Glist (n, 6, 6)``, ``Mlist_per_link (n, 4, 4)``, ``joint_limits (n, 2)``.  They were
extracted by running the unmodified reference (``oracle/gen_golden.py``); URDF parsing itself
is outside the hot path.
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Any, Optional

import numpy as np

_DIR = Path(__file__).resolve().parent


def available_robots() -> list[str]:
    return sorted(p.stem for p in _DIR.glob("*.npz") if not p.stem.endswith("_links"))


@dataclass
class RobotBundle:
    name: str
    S_list: np.ndarray
    M: np.ndarray
    Glist: np.ndarray
    Mlist_per_link: np.ndarray
    joint_limits: np.ndarray
    device: Optional[Any] = None
    _dyn: Any = None

    @property
    def num_joints(self) -> int:
        return int(self.S_list.shape[1])

    @property
    def dynamics(self):
        from ..dynamics import ManipulatorDynamics

        if self._dyn is None:
            self._dyn = ManipulatorDynamics(self.M, None, None, None, self.S_list, None, self.Glist,
                                            self.Mlist_per_link, device=self.device)
            # the reference hands the URDF's joint limits to its SerialManipulator (urdf/core.py:794):
            # the inverse-kinematics front ends clip to them and draw their guesses inside them
            self._dyn.joint_limits = [(float(lo), float(hi)) for lo, hi in np.asarray(self.joint_limits)]
        return self._dyn

    @property
    def serial_manipulator(self):
        return self.dynamics

    @property
    def links(self) -> dict:
        """Link table of the robot's URDF (names, the joint each link hangs on, home poses,
        allowed-collision matrix), extracted by the reference's loader (``<name>_links.npz``)."""
        path = _DIR / f"{self.name}_links.npz"
        if not path.exists():
            raise KeyError(f"no link table bundled for '{self.name}'")
        with np.load(path) as d:
            return {k: d[k] for k in d.files}

    def collision_checker(self, hulls):
        """``CollisionChecker`` over ``hulls = {link name: (V, 3) points in the link frame}``."""
        from ..potential_field import CollisionChecker

        return CollisionChecker(self.dynamics, self.links, hulls, device=self.device)

    def planner(self, torque_limits=None, **kw):
        from ..path_planning import OptimizedTrajectoryPlanning

        return OptimizedTrajectoryPlanning(self.dynamics, None, self.dynamics, self.joint_limits,
                                           torque_limits, device=self.device, **kw)


def load_robot(name: str, device: Optional[Any] = None) -> RobotBundle:
    path = _DIR / f"{name}.npz"
    if not path.exists():
        raise KeyError(f"unknown robot '{name}'. Available: {', '.join(available_robots())}")
    with np.load(path) as d:
        return RobotBundle(name, d["S_list"], d["M"], d["Glist"], d["Mlist_per_link"], d["joint_limits"],
                           device=device)
