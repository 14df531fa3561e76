"""In-tree build of the native libraries (no JIT cache: the built .so files travel to the GPU box).

``libmpk.so``      hand-written sm_100a kernels + the C ABI of ``include/mpk.h`` (nvcc, static cudart)
``_mpk_ops.so``    PyTorch custom-op extension (``torch.ops.mpk.*``) -- a thin caller of the C ABI

Run ``python -m manipulapy_b200._build`` (or ``__graft_entry__.build()``).
"""

from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "_lib" / "obj"
LIB = PKG / "_lib" / "libmpk.so"
OPS = PKG / "_lib" / "_mpk_ops.so"
INCLUDE = PKG.parent / "include"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]
# (source, object stem, extra defines).  The dynamics kernels exist in three flavours (rigid +
# all-revolute, rigid, general inertias; csrc/dyn_kernels.cuh), each its own translation unit
# so that the build uses every core.
CU_UNITS = [("robot.cu", "robot", []), ("traj.cu", "traj", os.environ.get("MPK_TRAJ_DEFINES", "").split()), ("kin.cu", "kin", []), ("ik.cu", "ik", []),
            ("dyn.cu", "dyn", []), ("peer.cu", "peer", []), ("collision.cu", "collision", []), ("legacy.cu", "legacy", [])]
# MPK_FD_DEFINES / MPK_DYN_DEFINES (environment, e.g. "-DMPK_FD_MINBLOCKS=16"): extra defines for
# the forward-dynamics / inverse-dynamics flavour units, for tuning sweeps on the GPU box.
CU_UNITS += [(f"{base}_flavour.cu", f"{base}_flavour{k}",
              [f"-DMPK_FLAVOUR={k}", *os.environ.get(f"MPK_{base.upper()}_DEFINES", "").split()])
             for base in ("dyn", "fd") for k in (0, 1, 2)]
# one translation unit pair per link-geometry signature that has its own kernels (MPK_GEO_LIST)
import re as _re

GEO_LIST = [(int(n), g) for n, g in _re.findall(r"X\((\d+), (0x[0-9a-f]+)u\)", (CSRC / "mpk_common.cuh").read_text())]
CU_UNITS += [(f"{base}_geo.cu", f"{base}_geo_{n}_{g[2:]}", [f"-DMPK_GEO_N={n}", f"-DMPK_GEO_SIG={g}u",
                                                        *os.environ.get(f"MPK_{base.upper()}_DEFINES", "").split()])
             for base in ("dyn", "fd") for n, g in GEO_LIST]
HEADERS = [CSRC / "mpk_device.cuh", CSRC / "mpk_common.cuh", CSRC / "dyn_kernels.cuh", INCLUDE / "mpk.h"]


def _digest(paths, extra="") -> str:
    h = hashlib.sha256(extra.encode())
    for p in paths:
        h.update(Path(p).read_bytes())
    return h.hexdigest()[:16]


def _stale(target: Path, stamp: str) -> bool:
    s = target.with_suffix(target.suffix + ".stamp")
    return not target.exists() or not s.exists() or s.read_text() != stamp


def _mark(target: Path, stamp: str) -> None:
    target.with_suffix(target.suffix + ".stamp").write_text(stamp)


def _run(cmd, log: Path | None = None) -> None:
    r = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if log is not None:
        log.write_text(r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {' '.join(map(str, cmd))}")


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu for sm_100a and link ``libmpk.so``."""
    OBJ.mkdir(parents=True, exist_ok=True)
    flags = " ".join(NVCC_FLAGS)

    def compile_one(unit) -> bool:
        name, stem, defines = unit
        src = CSRC / name
        obj = OBJ / (stem + ".o")
        stamp = _digest([src, *HEADERS], flags + " ".join(defines))
        if not force and not _stale(obj, stamp):
            return False
        if verbose:
            print(f"[mpk build] nvcc {name} {' '.join(defines)}", flush=True)
        _run([NVCC, *NVCC_FLAGS, *defines, "-I", INCLUDE, "-c", src, "-o", obj], OBJ / (stem + ".ptxas.log"))
        _mark(obj, stamp)
        return True

    # the heavy flavour units first so that they overlap with each other
    units = sorted(CU_UNITS, key=lambda u: "flavour" not in u[0] and "geo" not in u[0])
    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 1)) as ex:
        rebuilt = list(ex.map(compile_one, units))
    if any(rebuilt) or not LIB.exists():
        if verbose:
            print("[mpk build] link libmpk.so", flush=True)
        _run([NVCC, "-shared", "-o", LIB, *[OBJ / (stem + ".o") for _, stem, _ in CU_UNITS],
              "-cudart", "static"])
    return LIB


def build_ops(force: bool = False, verbose: bool = False) -> Path:
    """Compile the torch custom-op extension against the installed torch and link it to libmpk.so."""
    import torch
    from torch.utils import cpp_extension as ce

    src = CSRC / "mpk_torch.cpp"
    stamp = _digest([src, INCLUDE / "mpk.h"], torch.__version__)
    if not force and not _stale(OPS, stamp):
        return OPS
    if verbose:
        print("[mpk build] g++ mpk_torch.cpp", flush=True)
    inc = []
    for p in ce.include_paths():
        inc += ["-isystem", p]
    torch_lib = Path(torch.__file__).parent / "lib"
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
           "-DTORCH_EXTENSION_NAME=_mpk_ops", "-I", INCLUDE, "-I", "/usr/local/cuda/include", *inc,
           src, "-o", OPS, f"-L{LIB.parent}", "-l:libmpk.so", f"-L{torch_lib}", "-lc10", "-ltorch_cpu",
           "-ltorch", "-lc10_cuda", "-ltorch_cuda", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}"]
    _run(cmd)
    _mark(OPS, stamp)
    return OPS


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_lib(force, verbose)
    build_ops(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    for log in sorted(OBJ.glob("*.ptxas.log")):
        spills = [l for l in log.read_text().splitlines() if "spill" in l and "0 bytes spill stores, 0 bytes spill loads" not in l]
        print(f"{log.name}: {len(spills)} kernels with spills")
