"""manipulapy_b200 -- B200-native (sm_100a) batched trajectory-and-dynamics hot path of
ManipulaPy behind the reference's own Python API.

    from manipulapy_b200 import load_robot
    robot = load_robot("ur5")                       # bundled constant pack (from the reference URDF loader)
    planner = robot.planner()
    traj = planner.batch_joint_trajectory(start, end, Tf=2.0, N=2441, method=5)
    tau = planner.inverse_dynamics_trajectory(traj["positions"], traj["velocities"], traj["accelerations"])

The kernels live in ``csrc/`` behind the C ABI of ``include/mpk.h``; PyTorch provides
device memory, streams and ``torch.distributed``.  There is no CPU fallback.
"""

from .cuda_kernels import KERNEL_REGISTRY, KernelRegistration, KernelRegistry, execute_registered_kernel
from .dynamics import ManipulatorDynamics
from .kinematics import SerialManipulator
from .path_planning import OptimizedTrajectoryPlanning, TrajectoryPlanning
from .potential_field import CollisionChecker, PotentialField
from .robots import RobotBundle, available_robots, load_robot
from .sharding import gather_rows, shard_range
from .singularity import Singularity
from ._host import bind_host_to_device
from . import ik_helpers

__version__ = "0.1.0"

__all__ = [
    "KERNEL_REGISTRY", "KernelRegistration", "KernelRegistry", "execute_registered_kernel",
    "ManipulatorDynamics", "SerialManipulator", "OptimizedTrajectoryPlanning", "TrajectoryPlanning",
    "RobotBundle", "available_robots", "load_robot", "gather_rows", "shard_range", "bind_host_to_device",
    "ik_helpers", "Singularity", "CollisionChecker", "PotentialField",
]
