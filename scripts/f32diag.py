import numpy as np, torch, sys
sys.path.insert(0, '.')
from manipulapy_b200 import load_robot
for robot in ("ur5", "iiwa14", "panda"):
    rb = load_robot(robot); n = rb.num_joints
    rng = np.random.default_rng(9)
    B, N = 23, 257
    s, e = rng.uniform(-3, 3, (B, n)), rng.uniform(-3, 3, (B, n))
    planner = rb.planner(torque_limits=np.array([[-60.0, 55.0]] * n))
    for ft in (None, [1.0, -2.0, 0.5, 3.0, 0.0, -1.0]):
        tau32, tr = planner.trajectory_inverse_dynamics(s, e, 2.0, N, 5, [0, 0, -9.81], ft, return_trajectory=True, precision="float32")
        two32 = planner.inverse_dynamics_trajectory(tr["positions"], tr["velocities"], tr["accelerations"], [0, 0, -9.81], ft, precision="float32")
        d = np.abs(tau32.astype(np.float64) - two32)
        neq = tau32.view(np.uint32) != two32.view(np.uint32)
        idx = np.argwhere(neq)
        print(robot, ft is not None, "differing", int(neq.sum()), "of", neq.size, "max abs", d.max(), "rel", (d / np.maximum(1e-30, np.abs(two32))).max(), "first idx", idx[:5].tolist(), "by joint", neq.sum((0, 1)).tolist())
