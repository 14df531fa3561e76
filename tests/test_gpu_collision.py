"""Collision / limit post-processing hook of ``joint_trajectory`` on the GPU, against
what the UNMODIFIED reference did with the same injected hulls (tests/golden/collision.npz, generated
by oracle/gen_collision_golden.py) and against the numpy oracle on fresh random inputs.

Bars: link poses 1e-12 (float64 chain products in a different order than NumPy's); collision flags
bit-exact booleans; nudged rows bit-exact float32 (the nudge is float32 arithmetic with every
operation rounded, and the rows only depend on the flags)."""

import numpy as np
import pytest

from conftest import ROBOTS as ROBOT_DIR
from conftest import load_golden, load_pack

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _case(robot):
    from manipulapy_b200 import load_robot
    from oracle.oracle_lib import CollisionOracle

    g = load_golden("collision")
    rb = load_robot(robot)
    hulls = {str(nm): g[f"{robot}_hull_{nm}"] for nm in g[f"{robot}_hull_links"]}
    ck = rb.collision_checker(hulls)
    links = rb.links
    names = [str(x) for x in links["link_names"]]
    co = CollisionOracle(load_pack(robot)["S_list"], links["link_joint"], links["link_home"],
                         {names.index(k): v for k, v in hulls.items()}, links["link_acm"])
    return g, rb, ck, co, names


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_link_fk_and_flags_vs_reference_golden(robot):
    g, rb, ck, co, names = _case(robot)
    cfgs = g[f"{robot}_cfgs"]
    fk = ck.link_fk_batch(cfgs)
    assert list(fk) == names
    for l, nm in enumerate(names):
        np.testing.assert_allclose(fk[nm], g[f"{robot}_link_fk"][:, l], rtol=0, atol=1e-12)
    one = ck.link_fk(cfgs[3])
    np.testing.assert_allclose(one[names[-1]], g[f"{robot}_link_fk"][3, -1], rtol=0, atol=1e-12)
    flags = ck.check_collision(cfgs)
    assert flags.dtype == np.uint8 and np.array_equal(flags, g[f"{robot}_flags"])
    assert ck.check_collision(cfgs[0]) is bool(g[f"{robot}_flags"][0])
    # float32 rows (what the trajectory hook feeds the checker) and device tensors
    for k in (0, 1):
        t = f"{robot}_traj{k}_"
        assert np.array_equal(ck.check_collision(g[t + "raw"]), g[t + "flags_before"])
        dev = ck.check_collision(torch.from_numpy(g[t + "positions"]).cuda())
        assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), g[t + "flags_after"])


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_joint_trajectory_with_the_hook_vs_reference_golden(robot):
    """planner.joint_trajectory with a checker attached = the reference's joint_trajectory with the same
    hulls: every row, nudged or not, bit for bit; velocities / accelerations untouched."""
    g, rb, ck, co, names = _case(robot)
    planner = rb.planner()
    from manipulapy_b200 import PotentialField

    planner.attach_collision_checker(ck, potential_field=PotentialField(*g[f"{robot}_gains"]))
    for k in (0, 1):
        t = f"{robot}_traj{k}_"
        Tf, N, method = g[t + "args"]
        res = planner.joint_trajectory(g[t + "start"], g[t + "end"], float(Tf), int(N), int(method))
        for key in ("positions", "velocities", "accelerations"):
            assert res[key].dtype == np.float32
            assert np.array_equal(res[key].view(np.uint32), g[t + key].view(np.uint32)), (robot, k, key)
        moved = (res["positions"] != g[t + "raw"]).any(1)
        assert moved.sum() > 0 and not moved[g[t + "flags_before"] == 0].any()
        # the pieces: iteration counts and final flags
        rows, iters, still = ck.avoid(g[t + "raw"], g[t + "end"].astype(np.float32), planner.potential_field,
                                      return_info=True)
        assert np.array_equal(rows.view(np.uint32), g[t + "positions"].view(np.uint32))
        assert np.array_equal(still, g[t + "flags_after"])
        _, it_ref = co.avoid(g[t + "raw"][::7], g[t + "end"].astype(np.float32), g[f"{robot}_gains"][0])
        assert np.array_equal(iters[::7], it_ref)
        if k == 1:
            assert (iters == 100).any()
    # without a checker the hook is off and the rows are the raw ones
    plain = rb.planner().joint_trajectory(g[f"{robot}_traj0_start"], g[f"{robot}_traj0_end"], 2.0, 1000, 5)
    assert np.array_equal(plain["positions"].view(np.uint32), g[f"{robot}_traj0_raw"].view(np.uint32))


@pytest.mark.parametrize("robot", ["ur5", "iiwa14"])
def test_batched_checker_vs_oracle(robot):
    """Fresh random inputs, sizes past one block, a batch of trajectories with one goal each."""
    g, rb, ck, co, names = _case(robot)
    n = rb.num_joints
    rng = np.random.default_rng(31)
    lim = rb.joint_limits
    cfgs = rng.uniform(lim[:, 0], lim[:, 1], (333, n))
    assert np.array_equal(ck.check_collision(cfgs), co.check_collision(cfgs))
    np.testing.assert_allclose(np.stack(list(ck.link_fk_batch(cfgs[:40]).values()), 1), co.link_fk_batch(cfgs[:40]),
                               rtol=0, atol=1e-12)
    B, N = 5, 37
    planner = rb.planner()
    s, e = rng.uniform(lim[:, 0], lim[:, 1], (B, n)), rng.uniform(lim[:, 0], lim[:, 1], (B, n))
    raw = planner.batch_joint_trajectory(s, e, 1.0, N, 3)["positions"]
    got, iters, still = ck.avoid(raw, e.astype(np.float32), return_info=True)
    for b in range(B):
        ref, it_ref = co.avoid(raw[b], e[b].astype(np.float32))
        assert np.array_equal(got[b].view(np.uint32), ref.view(np.uint32))
        assert np.array_equal(iters.reshape(B, N)[b], it_ref)
    # a checker without hulls never reports a collision (the reference's state without meshes)
    empty = rb.collision_checker({})
    assert not empty.check_collision(cfgs).any() and empty.check_collision(cfgs[0]) is False
    assert np.array_equal(empty.avoid(raw, e.astype(np.float32)).view(np.uint32), raw.view(np.uint32))
