"""The drop-in boundary exercised the way INTEGRATION.md section B binds it: raw device pointers
(``tensor.data_ptr()``) and a ``cudaStream_t`` passed through ``ctypes`` into the ``extern "C"``
entry points of ``libmpk.so`` -- no ``torch.ops`` in between.  Results are compared bit for bit
with the ``torch.ops.mpk`` path and with the oracle; bad arguments must come back as status codes
with a reason in ``mpk_last_error()`` (nothing throws across the C boundary).

The launcher below is INTEGRATION.md's ``_launch_trajectory_mpk`` with torch in place of CuPy:
the signature of the reference's registry launchers (cuda_kernels/registry.py:828-867).
"""

import ctypes as C

import numpy as np
import pytest

from conftest import load_pack

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

MPK_OK, MPK_EINVAL, MPK_EUNSUPPORTED = 0, -1, -2
F64, F32 = 0, 1


@pytest.fixture(scope="module")
def lib():
    from manipulapy_b200 import _native

    assert torch.cuda.is_available()
    L = _native.lib()
    vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
    L.mpk_joint_trajectory.restype = C.c_int
    L.mpk_joint_trajectory.argtypes = [C.c_int, i64, i64, vp, vp, C.c_int, dbl, C.c_int, vp, vp, vp, vp, vp, vp]
    L.mpk_inverse_dynamics.restype = C.c_int
    L.mpk_inverse_dynamics.argtypes = [vp, i64, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp]
    L.mpk_forward_dynamics_trajectory.restype = C.c_int
    L.mpk_forward_dynamics_trajectory.argtypes = [vp, i64, i64, vp, vp, vp, C.c_int, vp, vp, dbl, C.c_int, vp,
                                                  vp, vp, vp, vp]
    L.mpk_trajectory_inverse_dynamics.restype = C.c_int
    L.mpk_trajectory_inverse_dynamics.argtypes = [vp, i64, i64, vp, vp, C.c_int, dbl, C.c_int, vp, vp, vp, vp,
                                                  vp, vp, vp, vp, vp, vp]
    L.mpk_fk_jacobian_space.restype = C.c_int
    L.mpk_fk_jacobian_space.argtypes = [vp, i64, vp, C.c_int, vp, vp, vp]
    L.mpk_mass_matrix.restype = C.c_int
    L.mpk_mass_matrix.argtypes = [vp, i64, vp, C.c_int, vp, vp]
    return L


def _handle(L, pack):
    h = C.c_void_p()
    S, M, G, Mc = (np.ascontiguousarray(pack[k], np.float64) for k in ("S_list", "M", "Glist", "Mlist_per_link"))
    rc = L.mpk_robot_create(S.shape[1], S.ctypes.data, M.ctypes.data, G.ctypes.data, Mc.ctypes.data, 0, C.byref(h))
    assert rc == MPK_OK, L.mpk_last_error()
    return h


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _hptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _launch_trajectory_mpk(L, thetastart, thetaend, Tf, N, method, limits=None):
    """(thetastart, thetaend, Tf, N, method) -> host float32 (pos, vel, acc)."""
    n = len(thetastart)
    s = torch.from_numpy(np.asarray(thetastart, np.float32).astype(np.float64)).cuda().reshape(1, n)
    e = torch.from_numpy(np.asarray(thetaend, np.float32).astype(np.float64)).cuda().reshape(1, n)
    pos, vel, acc = (torch.empty((1, N, n), dtype=torch.float32, device="cuda") for _ in range(3))
    lim = None if limits is None else np.ascontiguousarray(limits, np.float32)
    rc = L.mpk_joint_trajectory(n, 1, N, s.data_ptr(), e.data_ptr(), 1, float(Tf), int(method), _hptr(lim),
                                pos.data_ptr(), vel.data_ptr(), acc.data_ptr(), None, _stream())
    if rc != MPK_OK:
        raise RuntimeError(L.mpk_last_error().decode())
    return pos[0].cpu().numpy(), vel[0].cpu().numpy(), acc[0].cpu().numpy()


def _bits(a, b):
    return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a).view(np.uint32),
                                                 np.ascontiguousarray(b).view(np.uint32))


def test_joint_trajectory_through_ctypes(lib):
    from manipulapy_b200 import _native
    from oracle import Oracle

    pack = load_pack("ur5")
    rng = np.random.default_rng(11)
    lim32 = pack["joint_limits"].astype(np.float32)
    for N, method in ((1000, 5), (257, 3), (2, 5)):
        s, e = rng.uniform(-1, 1, 6), rng.uniform(-1, 1, 6)
        pos, vel, acc = _launch_trajectory_mpk(lib, s, e, 2.0, N, method, lim32)
        ref = Oracle.joint_trajectory(s.astype(np.float32)[None], e.astype(np.float32)[None], 2.0, N, method,
                                      pack["joint_limits"], inputs_f32=True)
        assert _bits(pos, ref["positions"][0]) and _bits(vel, ref["velocities"][0]) and _bits(acc, ref["accelerations"][0])
        # ... and the same bits as the torch.ops path
        sd = torch.from_numpy(s).cuda().reshape(1, 6)
        ed = torch.from_numpy(e).cuda().reshape(1, 6)
        p2, v2, a2 = _native.ops().joint_trajectory(sd, ed, True, 2.0, N, method, torch.from_numpy(lim32))
        assert _bits(pos, p2[0].cpu().numpy()) and _bits(vel, v2[0].cpu().numpy()) and _bits(acc, a2[0].cpu().numpy())


def test_inverse_dynamics_and_rollout_through_ctypes(lib):
    from manipulapy_b200 import _native, load_robot
    from oracle import Oracle

    ops = _native.ops()
    for name in ("ur5", "iiwa14"):
        pack = load_pack(name)
        n = pack["S_list"].shape[1]
        h = _handle(lib, pack)
        rb = load_robot(name)
        rng = np.random.default_rng(5)
        P = 1000
        th, dth, ddth = (torch.from_numpy(rng.uniform(-1, 1, (P, n))).cuda() for _ in range(3))
        g = (C.c_double * 3)(0.0, 0.0, -9.81)
        f = (C.c_double * 6)(1.0, -2.0, 0.5, 3.0, 1.5, -0.25)
        tl = np.ascontiguousarray(np.array([[-20.0, 20.0]] * n), np.float32)
        tau = torch.empty((P, n), dtype=torch.float32, device="cuda")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # a non-default stream handle crosses the boundary
            rc = lib.mpk_inverse_dynamics(h, P, th.data_ptr(), dth.data_ptr(), ddth.data_ptr(), F64, g, f, None,
                                          _hptr(tl), tau.data_ptr(), F32, _stream())
        assert rc == MPK_OK, lib.mpk_last_error()
        side.synchronize()
        ref_t = ops.inverse_dynamics(rb.dynamics.robot.handle, th, dth, ddth, [0.0, 0.0, -9.81],
                                     [1.0, -2.0, 0.5, 3.0, 1.5, -0.25], None, torch.from_numpy(tl), True, False)
        assert _bits(tau.cpu().numpy(), ref_t.cpu().numpy())
        orc = Oracle(pack["S_list"], pack["M"], pack["Glist"], pack["Mlist_per_link"])
        ref = orc.inverse_dynamics_trajectory(th.cpu().numpy(), dth.cpu().numpy(), ddth.cpu().numpy(),
                                              Ftip=np.array([1.0, -2.0, 0.5, 3.0, 1.5, -0.25]), torque_limits=tl,
                                              analytic=True)
        np.testing.assert_allclose(tau.cpu().numpy(), ref, rtol=3e-7, atol=1e-6)

        # forward-dynamics rollouts, B = 3, float32 torque rows
        B, N = 3, 40
        th0 = torch.from_numpy(rng.uniform(-0.5, 0.5, (B, n))).cuda()
        dth0 = torch.from_numpy(rng.uniform(-0.2, 0.2, (B, n))).cuda()
        taum = torch.from_numpy(rng.uniform(-3, 3, (B, N, n)).astype(np.float32)).cuda()
        jl = np.ascontiguousarray(pack["joint_limits"], np.float32)
        outs = [torch.empty((B, N, n), dtype=torch.float32, device="cuda") for _ in range(3)]
        rc = lib.mpk_forward_dynamics_trajectory(h, B, N, th0.data_ptr(), dth0.data_ptr(), taum.data_ptr(), F32, g,
                                                 None, 1e-3, 2, _hptr(jl), outs[0].data_ptr(), outs[1].data_ptr(),
                                                 outs[2].data_ptr(), _stream())
        assert rc == MPK_OK, lib.mpk_last_error()
        torch.cuda.synchronize()
        ref3 = ops.forward_dynamics_trajectory(rb.dynamics.robot.handle, th0, dth0, taum, [0.0, 0.0, -9.81], None,
                                               1e-3, 2, torch.from_numpy(jl))
        for a, b in zip(outs, ref3):
            assert _bits(a.cpu().numpy(), b.cpu().numpy())
        lib.mpk_robot_destroy(h)


def test_fused_and_kinematics_through_ctypes(lib):
    from manipulapy_b200 import _native, load_robot

    ops = _native.ops()
    pack = load_pack("ur5")
    h = _handle(lib, pack)
    rb = load_robot("ur5")
    rng = np.random.default_rng(9)
    B, N = 37, 301
    s = torch.from_numpy(rng.uniform(-3, 3, (B, 6))).cuda()
    e = torch.from_numpy(rng.uniform(-3, 3, (B, 6))).cuda()
    g = (C.c_double * 3)(0.0, 0.0, -9.81)
    jl = np.ascontiguousarray(pack["joint_limits"], np.float32)
    tau = torch.empty((B, N, 6), dtype=torch.float32, device="cuda")
    scratch = torch.empty((3, N), dtype=torch.float64, device="cuda")
    rc = lib.mpk_trajectory_inverse_dynamics(h, B, N, s.data_ptr(), e.data_ptr(), 0, 2.0, 5, _hptr(jl), g, None, None,
                                             tau.data_ptr(), None, None, None, scratch.data_ptr(), _stream())
    assert rc == MPK_OK, lib.mpk_last_error()
    ref = ops.trajectory_inverse_dynamics(rb.dynamics.robot.handle, s, e, False, 2.0, N, 5, torch.from_numpy(jl),
                                          [0.0, 0.0, -9.81], None, None, False, False)[0]
    torch.cuda.synchronize()
    assert _bits(tau.cpu().numpy(), ref.cpu().numpy())

    P = 999
    th = torch.from_numpy(rng.uniform(-3, 3, (P, 6))).cuda()
    T = torch.empty((P, 4, 4), dtype=torch.float64, device="cuda")
    J = torch.empty((P, 6, 6), dtype=torch.float64, device="cuda")
    Mm = torch.empty((P, 6, 6), dtype=torch.float64, device="cuda")
    assert lib.mpk_fk_jacobian_space(h, P, th.data_ptr(), F64, T.data_ptr(), J.data_ptr(), _stream()) == MPK_OK
    assert lib.mpk_mass_matrix(h, P, th.data_ptr(), F64, Mm.data_ptr(), _stream()) == MPK_OK
    torch.cuda.synchronize()
    T2, J2 = ops.fk_jacobian(rb.dynamics.robot.handle, th, True, True, False, False)
    assert torch.equal(T, T2) and torch.equal(J, J2)
    assert torch.equal(Mm, ops.mass_matrix(rb.dynamics.robot.handle, th))
    lib.mpk_robot_destroy(h)


def test_bad_arguments_return_status_codes(lib):
    pack = load_pack("ur5")
    h = _handle(lib, pack)
    g = (C.c_double * 3)(0.0, 0.0, -9.81)
    x = torch.zeros((8, 6), dtype=torch.float64, device="cuda")
    out = torch.zeros((8, 16, 6), dtype=torch.float32, device="cuda")
    st = _stream()

    # NULL pointers
    assert lib.mpk_joint_trajectory(6, 8, 16, None, x.data_ptr(), 0, 2.0, 5, None, out.data_ptr(), None, None, None,
                                    st) == MPK_EINVAL
    assert b"NULL" in lib.mpk_last_error() or b"required" in lib.mpk_last_error()
    assert lib.mpk_inverse_dynamics(None, 8, x.data_ptr(), None, None, F64, g, None, None, None, out.data_ptr(), F32,
                                    st) == MPK_EINVAL
    assert lib.mpk_inverse_dynamics(h, 8, None, None, None, F64, g, None, None, None, out.data_ptr(), F32,
                                    st) == MPK_EINVAL
    assert lib.mpk_forward_dynamics_trajectory(h, 8, 16, x.data_ptr(), x.data_ptr(), None, F32, g, None, 1e-3, 1, None,
                                               out.data_ptr(), out.data_ptr(), out.data_ptr(), st) == MPK_EINVAL
    # more joints than the kernels are instantiated for
    assert lib.mpk_joint_trajectory(9, 8, 16, x.data_ptr(), x.data_ptr(), 0, 2.0, 5, None, out.data_ptr(), None, None,
                                    None, st) == MPK_EUNSUPPORTED
    assert b"dof" in lib.mpk_last_error()
    S9 = np.zeros((6, 9))
    S9[2] = 1.0
    h9 = C.c_void_p()
    assert lib.mpk_robot_create(9, S9.ctypes.data, np.eye(4).ctypes.data, None, None, 0,
                                C.byref(h9)) == MPK_EUNSUPPORTED
    assert not h9.value
    # misaligned output rows (the stores are 16-byte vectors)
    raw = torch.zeros(8 * 16 * 6 + 4, dtype=torch.float32, device="cuda")
    assert lib.mpk_joint_trajectory(6, 8, 16, x.data_ptr(), x.data_ptr(), 0, 2.0, 5, None, raw.data_ptr() + 4, None,
                                    None, None, st) == MPK_EINVAL
    assert b"aligned" in lib.mpk_last_error()
    # bad dtype code, negative sizes, zero sub-steps
    assert lib.mpk_inverse_dynamics(h, 8, x.data_ptr(), None, None, 7, g, None, None, None, out.data_ptr(), F32,
                                    st) == MPK_EINVAL
    assert lib.mpk_joint_trajectory(6, -1, 16, x.data_ptr(), x.data_ptr(), 0, 2.0, 5, None, out.data_ptr(), None,
                                    None, None, st) == MPK_EINVAL
    assert lib.mpk_forward_dynamics_trajectory(h, 8, 16, x.data_ptr(), x.data_ptr(), out.data_ptr(), F32, g, None,
                                               1e-3, 0, None, out.data_ptr(), out.data_ptr(), out.data_ptr(),
                                               st) == MPK_EINVAL
    # empty batches are fine and touch nothing
    assert lib.mpk_joint_trajectory(6, 0, 16, None, None, 0, 2.0, 5, None, None, None, None, None, st) == MPK_OK
    # a kinematics-only robot refuses dynamics
    hk = C.c_void_p()
    S = np.ascontiguousarray(pack["S_list"], np.float64)
    M = np.ascontiguousarray(pack["M"], np.float64)
    assert lib.mpk_robot_create(6, S.ctypes.data, M.ctypes.data, None, None, 0, C.byref(hk)) == MPK_OK
    assert lib.mpk_inverse_dynamics(hk, 8, x.data_ptr(), None, None, F64, g, None, None, None, out.data_ptr(), F32,
                                    st) == MPK_EINVAL
    assert b"Glist" in lib.mpk_last_error()
    lib.mpk_robot_destroy(hk)
    lib.mpk_robot_destroy(h)
    torch.cuda.synchronize()


def test_fused_kernel_stores_into_row_offset_views():
    """The multi-GPU gather has every rank's kernel store its rows into a slice of the collecting
    rank's buffer (sharding.PeerRows): the destination's first byte is then only 8- or 4-byte aligned
    (2441 x 6 x 4 B per trajectory is not a multiple of 16).  Same bits as a fresh 16-byte-aligned result."""
    from manipulapy_b200 import _native, load_robot
    from manipulapy_b200.sharding import PeerRows

    ops = _native.ops()
    rb = load_robot("ur5")
    h, jl = rb.dynamics.robot.handle, rb.planner()._jl
    rng = np.random.default_rng(21)
    B, N = 9, 2441
    s = torch.from_numpy(rng.uniform(-3, 3, (B, 6))).cuda()
    e = torch.from_numpy(rng.uniform(-3, 3, (B, 6))).cuda()
    g = [0.0, 0.0, -9.81]
    ref = ops.trajectory_inverse_dynamics(h, s, e, False, 2.0, N, 5, jl, g, None, None, False)[0]
    big = torch.full((B + 2, N, 6), float("nan"), dtype=torch.float32, device="cuda")
    for lo in (0, 1, 2, 3):  # byte offsets 0, 8 (mod 16), 0, 8
        ops.trajectory_inverse_dynamics(h, s[lo:], e[lo:], False, 2.0, N, 5, jl, g, None, None, False, False, big[lo:B])
        assert torch.equal(big[lo:B], ref[lo:]) and bool(torch.isnan(big[B:]).all())
    # 4-byte-aligned destination: a flat buffer entered one float in
    flat = torch.full((B * N * 6 + 1,), float("nan"), dtype=torch.float32, device="cuda")
    ops.trajectory_inverse_dynamics(h, s, e, False, 2.0, N, 5, jl, g, None, None, False, False, flat[1:])
    assert torch.equal(flat[1:].view(B, N, 6), ref)
    # the single-rank PeerRows is the same code path with an ordinary buffer
    pr = PeerRows(B, (N, 6), torch.float32, torch.device("cuda", torch.cuda.current_device()))
    ops.trajectory_inverse_dynamics(h, s, e, False, 2.0, N, 5, jl, g, None, None, False, False, pr.rows())
    pr.commit()
    assert torch.equal(pr.full, ref)
    pr.close()
