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


def load_pack(name: str) -> dict:
    with np.load(ROBOTS / f"{name}.npz") as d:
        return {k: d[k] for k in d.files}


def load_golden(name: str) -> dict:
    with np.load(GOLDEN / f"{name}.npz") as d:
        return {k: d[k] for k in d.files}


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
