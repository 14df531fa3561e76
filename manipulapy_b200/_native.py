"""Loader of the native libraries.  Fails loudly: there is no Python / CPU fallback.

* ``lib()``  -- ``libmpk.so`` through ctypes (the C ABI of ``include/mpk.h``).  Host-only
  entry points (robot pack construction, version, error string) work without a GPU.
* ``ops()``  -- ``torch.ops.mpk`` after loading ``_mpk_ops.so`` (the PyTorch custom-op
  extension, linked against ``libmpk.so``).
"""

from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import os

PKG = Path(__file__).resolve().parent
# MPK_LIB_DIR (tuning sweeps only): load a differently compiled pair of libraries, e.g. one built with
# other launch bounds (scripts/build_variant.sh)
_LIB_DIR = Path(os.environ["MPK_LIB_DIR"]).resolve() if os.environ.get("MPK_LIB_DIR") else PKG / "_lib"
LIB_PATH = _LIB_DIR / "libmpk.so"
OPS_PATH = _LIB_DIR / "_mpk_ops.so"
HEADER = PKG.parent / "include" / "mpk.h"

_lib = None
_ops = None


class NativeLibraryMissing(ImportError):
    pass


def _require(path: Path) -> Path:
    if not path.exists():
        raise NativeLibraryMissing(
            f"{path} is missing: build it with `python -m manipulapy_b200._build` "
            "(nvcc, sm_100a).  manipulapy_b200 has no CPU fallback."
        )
    return path


def declared_symbols() -> list[str]:
    """Every function ``include/mpk.h`` declares (used by the export test)."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpk_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(str(_require(LIB_PATH)))
        L.mpk_version.restype = C.c_int
        L.mpk_last_error.restype = C.c_char_p
        L.mpk_robot_create.restype = C.c_int
        L.mpk_robot_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.POINTER(C.c_void_p)]
        L.mpk_robot_destroy.restype = None
        L.mpk_robot_destroy.argtypes = [C.c_void_p]
        L.mpk_robot_dof.argtypes = [C.c_void_p]
        L.mpk_robot_is_rigid.argtypes = [C.c_void_p]
        L.mpk_robot_all_revolute.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def ops():
    global _ops
    if _ops is None:
        import torch

        lib()  # resolve libmpk.so first so the extension's DT_NEEDED entry is satisfied
        torch.ops.load_library(str(_require(OPS_PATH)))
        _ops = torch.ops.mpk
    return _ops


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError(
            "manipulapy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback. "
            "Use the reference ManipulaPy for CPU execution."
        )
