"""Shared fixtures.  ``-m "not gpu"`` runs everywhere; ``-m gpu`` needs a B200."""

import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden"
ROBOTS = REPO / "manipulapy_b200" / "robots"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _native_libraries():
    """A fresh checkout has no built libraries (they are git-ignored): build them once, in-tree,
    like ``__graft_entry__.build()`` does.  Nothing is rebuilt when they are there."""
    from manipulapy_b200 import _build, _native

    if not (_native.LIB_PATH.exists() and _native.OPS_PATH.exists()):
        _build.build_all()


def load_pack(name: str) -> dict:
    with np.load(ROBOTS / f"{name}.npz") as d:
        return {k: d[k] for k in d.files}


def load_golden(name: str) -> dict:
    with np.load(GOLDEN / f"{name}.npz") as d:
        return {k: d[k] for k in d.files}


def load_zoo() -> dict:
    """Robots of the reference's whole bundled database (oracle/gen_robot_zoo.py): name -> dict
    with the constant pack and the reference's outputs at a few configurations."""
    with np.load(GOLDEN / "robot_zoo.npz") as d:
        zoo = {str(r): {} for r in d["robots"]}
        for k in d.files:
            if "/" in k:
                r, f = k.split("/", 1)
                zoo[r][f] = d[k]
        for r in zoo:
            zoo[r]["g"] = d["g"]
    return zoo


ZOO_ROBOTS = sorted(load_zoo())


def _zoo_rel(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    sc = np.maximum(1.0, np.abs(ref).reshape(ref.shape[0], -1).max(1))
    return float((np.abs(np.asarray(got) - ref).reshape(ref.shape[0], -1).max(1) / sc).max())


def check_zoo_outputs(z, fk, jac, M, grav, cor, tau, dd):
    """The reference's own golden tolerances (tests/test_dynamics_golden.py:77-83) for M, g, c and
    inverse dynamics; 1e-13 for FK / Jacobian; forward dynamics per vector."""
    np.testing.assert_allclose(fk, z["fk"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(jac, z["jac"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(M, z["mass"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(grav, z["g_forces"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(cor, z["c"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(tau, z["id"], rtol=1e-7, atol=1e-8)
    # the reference's finite-difference Coriolis noise (~1e-9 abs) is amplified by cond(M)
    # (up to ~1e5 for the grippers' gram-scale links), hence per vector at 1e-6
    assert _zoo_rel(dd, z["fd"]) < 1e-6


@pytest.fixture(scope="session")
def oracle_factory():
    from oracle import Oracle

    cache = {}

    def make(name: str):
        if name not in cache:
            p = load_pack(name)
            cache[name] = Oracle(p["S_list"], p["M"], p["Glist"], p["Mlist_per_link"])
        return cache[name]

    return make


def planar_2r_pack(L1=1.0, L2=1.0, m1=1.0, m2=1.0) -> dict:
    """2R planar arm with point masses (reference tests/test_v132_regressions.py:126-192)."""
    S = np.zeros((6, 2))
    S[2, :] = 1.0
    # v = -w x r, r1 = 0, r2 = (L1, 0, 0)
    S[3:, 1] = -np.cross([0, 0, 1.0], [L1, 0, 0])
    M = np.eye(4)
    M[0, 3] = L1 + L2
    Mc = np.stack([np.eye(4), np.eye(4)])
    Mc[0, 0, 3] = L1
    Mc[1, 0, 3] = L1 + L2
    G = np.stack([np.diag([0, 0, 0, m, m, m]) for m in (m1, m2)]).astype(float)
    lim = np.array([[-np.pi, np.pi]] * 2)
    return dict(S_list=S, M=M, Glist=G, Mlist_per_link=Mc, joint_limits=lim)


# ---- native library helpers ----------------------------------------------------------------
import ctypes as _C
import subprocess as _sp


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_C.c_void_p)


class HostCheck:
    """The kernels' own per-thread templates executed on the CPU (tests/hostcheck/hostcheck.cu).

    Lets the GPU-less container verify the CUDA path's algebra against the oracle and the
    reference's golden vectors.  Test infrastructure only."""

    def __init__(self):
        from manipulapy_b200 import _native

        here = REPO / "tests" / "hostcheck"
        src, so = here / "hostcheck.cu", here / "libhostcheck.so"
        deps = [src, REPO / "manipulapy_b200" / "csrc" / "mpk_device.cuh",
                REPO / "manipulapy_b200" / "csrc" / "mpk_common.cuh"]
        if not so.exists() or so.stat().st_mtime < max(p.stat().st_mtime for p in deps):
            _sp.run(["nvcc", "-O1", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
                     "-Wno-deprecated-gpu-targets", "--expt-relaxed-constexpr", "-I", str(REPO / "include"),
                     str(src), "-o", str(so), f"-L{_native.LIB_PATH.parent}", "-l:libmpk.so",
                     "-Xlinker", "-rpath", "-Xlinker", str(_native.LIB_PATH.parent)], check=True)
        self.L = _native.lib()
        self.H = _C.CDLL(str(so))

    def robot(self, pack, flags=0):
        h = _C.c_void_p()
        S = np.ascontiguousarray(pack["S_list"], dtype=np.float64)
        rc = self.L.mpk_robot_create(S.shape[1], _ptr(S), _ptr(np.ascontiguousarray(pack["M"], dtype=np.float64)),
                                     _ptr(np.ascontiguousarray(pack["Glist"], dtype=np.float64)),
                                     _ptr(np.ascontiguousarray(pack["Mlist_per_link"], dtype=np.float64)),
                                     flags, _C.byref(h))
        if rc != 0:
            raise RuntimeError(self.L.mpk_last_error().decode())
        return h, S.shape[1]

    def geo(self, rb):
        """Link-geometry signature of the robot (0: general kernels)."""
        self.L.mpk_robot_geometry_signature.restype = _C.c_uint
        return int(self.L.mpk_robot_geometry_signature(rb[0]))

    def set_geo(self, rb, geo):
        self.H.hc_set_geo.restype = _C.c_uint
        return int(self.H.hc_set_geo(rb[0], _C.c_uint(geo)))

    def rnea(self, rb, th, dth=None, ddth=None, g=(0, 0, -9.81), ftip=None, smem_store=False, f32=False):
        h, n = rb
        fn = self.H.hc_rnea_f32 if f32 else {0: self.H.hc_rnea, 1: self.H.hc_rnea_smem}[int(smem_store)]
        th = np.ascontiguousarray(th, dtype=np.float64).reshape(-1, n)
        P = th.shape[0]
        cv = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64).reshape(P, n)
        dth, ddth = cv(dth), cv(ddth)
        g = np.ascontiguousarray(g, dtype=np.float64)
        out = np.empty((P, n))
        if ftip is not None and np.ndim(ftip) == 2:
            for i in range(P):
                f = np.ascontiguousarray(ftip[i], dtype=np.float64)
                fn(h, _C.c_int64(1), _ptr(th[i:i + 1]), _ptr(None if dth is None else dth[i:i + 1]),
                               _ptr(None if ddth is None else ddth[i:i + 1]), _ptr(g),
                               _ptr(f) if f.any() else None, _ptr(out[i:i + 1]))
            return out
        f = None if ftip is None else np.ascontiguousarray(ftip, dtype=np.float64)
        fn(h, _C.c_int64(P), _ptr(th), _ptr(dth), _ptr(ddth), _ptr(g), _ptr(f), _ptr(out))
        return out

    def mass(self, rb, th):
        h, n = rb
        th = np.ascontiguousarray(th, dtype=np.float64).reshape(-1, n)
        out = np.empty((th.shape[0], n, n))
        self.H.hc_mass(h, _C.c_int64(th.shape[0]), _ptr(th), _ptr(out))
        return out

    def fk(self, rb, th, f32=False, body=False):
        h, n = rb
        th = np.ascontiguousarray(th, dtype=np.float64).reshape(-1, n)
        T, J = np.empty((th.shape[0], 4, 4)), np.empty((th.shape[0], 6, n))
        fn = self.H.hc_fk_body if body else (self.H.hc_fk_f32 if f32 else self.H.hc_fk)
        fn(h, _C.c_int64(th.shape[0]), _ptr(th), _ptr(T), _ptr(J))
        return T, J

    def kin_robot(self, S, M):
        """Kinematics-only handle."""
        h = _C.c_void_p()
        S = np.ascontiguousarray(S, dtype=np.float64)
        rc = self.L.mpk_robot_create(S.shape[1], _ptr(S), _ptr(np.ascontiguousarray(M, dtype=np.float64)), None, None,
                                     0, _C.byref(h))
        if rc != 0:
            raise RuntimeError(self.L.mpk_last_error().decode())
        return h, S.shape[1]

    def fd(self, rb, th, dth, tau, g=(0, 0, -9.81), ftip_rows=None):
        h, n = rb
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64).reshape(-1, n)
        th, dth, tau = c(th), c(dth), c(tau)
        f = None if ftip_rows is None else np.ascontiguousarray(ftip_rows, dtype=np.float64).reshape(-1, 6)
        out = np.empty_like(th)
        self.H.hc_fd(h, _C.c_int64(th.shape[0]), _ptr(th), _ptr(dth), _ptr(tau),
                     _ptr(np.ascontiguousarray(g, dtype=np.float64)), _ptr(f), _ptr(out))
        return out

    def ik(self, rb, Td, th0, eomg=1e-6, ev=1e-6, max_iterations=10000, damping=2e-2, step_cap=0.3,
           weight_orientation=1.0, weight_position=1.0, limits=None, seed=0, adaptive_tuning=False,
           backtracking=False, numpy_restarts=False):
        h, n = rb
        Td = np.ascontiguousarray(Td, dtype=np.float64).reshape(-1, 4, 4)
        th0 = np.ascontiguousarray(th0, dtype=np.float64).reshape(-1, n)
        P = Td.shape[0]
        lim = None if limits is None else np.ascontiguousarray(limits, dtype=np.float64).reshape(n, 2)
        th, it, ok = np.empty((P, n)), np.empty(P, np.int32), np.empty(P, np.uint8)
        if numpy_restarts:
            # the Python mirror's protocol for single-target calls (kinematics.py): restart noise drawn
            # from NumPy's global generator, which is then left where the reference would leave it
            from manipulapy_b200.kinematics import restart_noise_table, settle_generator

            assert P == 1
            noise, state = restart_noise_table(n, int(max_iterations))
            rs = np.zeros(1, np.int32)
            self.H.hc_ik(h, _C.c_int64(P), _ptr(Td), _ptr(th0), _C.c_double(eomg), _C.c_double(ev),
                         int(max_iterations), _C.c_double(damping), _C.c_double(step_cap),
                         _C.c_double(weight_orientation), _C.c_double(weight_position), _ptr(lim), _C.c_uint64(seed),
                         _ptr(th), _ptr(it), _ptr(ok), (1 if adaptive_tuning else 0) | (2 if backtracking else 0),
                         _ptr(noise), int(noise.shape[0]), _ptr(rs))
            settle_generator(state, n, int(rs[0]))
            return th, ok.astype(bool), it
        self.H.hc_ik(h, _C.c_int64(P), _ptr(Td), _ptr(th0), _C.c_double(eomg), _C.c_double(ev), int(max_iterations),
                     _C.c_double(damping), _C.c_double(step_cap), _C.c_double(weight_orientation),
                     _C.c_double(weight_position), _ptr(lim), _C.c_uint64(seed), _ptr(th), _ptr(it), _ptr(ok),
                     (1 if adaptive_tuning else 0) | (2 if backtracking else 0), None, 0, None)
        return th, ok.astype(bool), it

    def cartesian(self, Xs, Xe, Tf, N, method):
        Xs = np.ascontiguousarray(Xs, dtype=np.float64).reshape(4, 4)
        Xe = np.ascontiguousarray(Xe, dtype=np.float64).reshape(4, 4)
        pos, vel, acc = (np.empty((N, 3), np.float32) for _ in range(3))
        ori = np.empty((N, 3, 3), np.float32)
        self.H.hc_cartesian(_C.c_int64(N), _ptr(Xs), _ptr(Xe), _C.c_double(Tf), int(method), _ptr(pos), _ptr(vel),
                            _ptr(acc), _ptr(ori))
        return {"positions": pos, "velocities": vel, "accelerations": acc, "orientations": ori}

    def sincos(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).ravel()
        s, c = np.empty_like(x), np.empty_like(x)
        self.H.hc_sincos(_C.c_int64(x.size), _ptr(x), _ptr(s), _ptr(c))
        return s, c

    def traj(self, start, end, Tf, N, method, limits=None, inputs_f32=True):
        s = np.ascontiguousarray(start, dtype=np.float64)
        e = np.ascontiguousarray(end, dtype=np.float64)
        n = s.shape[0]
        lim = None if limits is None else np.ascontiguousarray(limits, dtype=np.float32)
        out = [np.empty((N, n), np.float32) for _ in range(3)]
        self.H.hc_traj(n, _C.c_int64(N), _ptr(s), _ptr(e), int(inputs_f32), _C.c_double(Tf), int(method),
                       _ptr(lim), *[_ptr(o) for o in out])
        return out


@pytest.fixture(scope="session")
def hostcheck():
    return HostCheck()


def random_general_pack(n, seed):
    """A chain with arbitrary unit screws / one prismatic joint, arbitrary CoM frames and full
    symmetric positive-definite 6x6 inertias (not rigid-body structured)."""
    rng = np.random.default_rng(seed)
    S = np.zeros((6, n))
    for i in range(n):
        w = rng.normal(size=3)
        w /= np.linalg.norm(w)
        q = rng.uniform(-0.5, 0.5, 3)
        if i == n // 2:
            S[3:, i] = w * 0.7  # prismatic with a non-unit direction
        else:
            S[:3, i] = w
            S[3:, i] = -np.cross(w, q)
    def rand_T():
        A = rng.normal(size=(3, 3))
        Q, _ = np.linalg.qr(A)
        if np.linalg.det(Q) < 0:
            Q[:, 0] *= -1
        T = np.eye(4)
        T[:3, :3] = Q
        T[:3, 3] = rng.uniform(-0.5, 0.5, 3)
        return T
    M = rand_T()
    Mc = np.stack([rand_T() for _ in range(n)])
    G = []
    for _ in range(n):
        A = rng.normal(size=(6, 6))
        G.append(A @ A.T + 0.5 * np.eye(6))
    lim = np.array([[-np.pi, np.pi]] * n)
    return dict(S_list=S, M=M, Glist=np.stack(G), Mlist_per_link=Mc, joint_limits=lim)
