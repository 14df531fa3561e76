"""Host <-> device plumbing shared by the API mirrors.

Callers pass host arrays (NumPy, lists) and get fresh NumPy arrays back, exactly like the
reference.  As an extension, ``torch`` CUDA tensors are accepted and then
results stay on the device as ``torch`` tensors (no copies, current stream).
"""

from __future__ import annotations

from typing import Any, Optional, Sequence

import numpy as np
import torch

from . import _native


def is_device_tensor(x: Any) -> bool:
    return isinstance(x, torch.Tensor) and x.is_cuda


def any_device(*xs: Any) -> bool:
    return any(is_device_tensor(x) for x in xs)


def default_device(device: Optional[Any] = None) -> torch.device:
    _native.require_cuda()
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    d = torch.device(device)
    if d.type != "cuda":
        raise RuntimeError("manipulapy_b200 runs on CUDA devices only (no CPU fallback)")
    return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())


def to_device(x: Any, device: torch.device, dtype: Optional[torch.dtype] = torch.float64,
              keep_f32: bool = False) -> torch.Tensor:
    """Host array / tensor -> contiguous CUDA tensor.

    ``keep_f32``: float32 inputs stay float32 in HBM (the kernels upcast exactly in
    registers); everything else becomes ``dtype``.
    """
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if not (a.dtype == np.float32 and keep_f32) and a.dtype != np.float64:
            a = a.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if keep_f32 and t.dtype == torch.float32:
        want = torch.float32
    else:
        want = dtype if dtype is not None else t.dtype
    if t.is_cuda:
        return t.to(device=device, dtype=want).contiguous()
    if t.numel() >= (1 << 16):
        # large host arrays: stage through pinned memory so the copy runs at PCIe speed
        pinned = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        pinned.copy_(t)
        return pinned.to(device, non_blocking=True).to(want).contiguous()
    return t.to(device).to(want).contiguous()


def promote_rows(*ts: Optional[torch.Tensor]):
    """Common storage type of the row inputs of one call: float32 only when EVERY row array is
    float32 (then the rows stay float32 in HBM and are upcast exactly in registers); any float64
    input promotes all of them to float64, as NumPy promotion does in the reference."""
    live = [t for t in ts if t is not None]
    want = torch.float32 if live and all(t.dtype == torch.float32 for t in live) else torch.float64
    return tuple(None if t is None else (t if t.dtype == want else t.to(want)) for t in ts)


def to_host(t: torch.Tensor) -> np.ndarray:
    """CUDA tensor -> fresh NumPy array (through pinned memory for large results)."""
    if t.numel() >= (1 << 16):
        out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        out.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return out.numpy()
    return t.cpu().numpy()


_copy_streams: dict = {}


def _copy_stream(device: torch.device) -> torch.cuda.Stream:
    key = (device.type, device.index)
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=device)
    return _copy_streams[key]


def host_out(out: Any, shape: Sequence[int], dtype: torch.dtype) -> torch.Tensor:
    """Caller-owned host destination (``out=`` of the API mirrors): a NumPy array or a CPU torch
    tensor of the result's shape and dtype, C-contiguous.  Pinned memory (``torch.empty(...,
    pin_memory=True)``, or its ``.numpy()`` view) receives the rows at PCIe speed; pageable memory
    works too, through the driver's staging copy."""
    t = out if isinstance(out, torch.Tensor) else torch.from_numpy(out)
    if t.is_cuda or t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
        raise ValueError(f"out must be a C-contiguous host array of shape {tuple(shape)} and dtype {dtype}")
    return t


def chunked_to_host(launch, rows: int, tail: Sequence[int], dtype: torch.dtype, device: torch.device,
                    chunks: int = 8, min_rows: int = 64, out: Any = None):
    """Run ``launch(lo, hi) -> CUDA tensor (hi - lo, *tail)`` over contiguous row chunks and
    stream every finished chunk to pinned host memory on a second stream, so the PCIe copy of
    chunk k overlaps the kernel of chunk k + 1.  Returns a fresh host array ``(rows, *tail)`` --
    or fills and returns the caller's ``out`` (no pinned allocation per call: a caller that keeps
    its results would otherwise pin host memory without bound, ~100 ms of ``cudaHostAlloc`` each)."""
    ret = None
    if out is not None:
        ret, out = out, host_out(out, (rows, *tail), dtype)
    else:
        out = torch.empty((rows, *tail), dtype=dtype, pin_memory=True)
    if rows == 0:
        return ret if ret is not None else out.numpy()
    per = max(min_rows, -(-rows // max(1, chunks)))
    compute = torch.cuda.current_stream(device)
    side = _copy_stream(device)
    keep = []
    for lo in range(0, rows, per):
        hi = min(rows, lo + per)
        t = launch(lo, hi)
        done = torch.cuda.Event()
        done.record(compute)
        side.wait_event(done)
        with torch.cuda.stream(side):
            out[lo:hi].copy_(t, non_blocking=True)
        t.record_stream(side)
        keep.append(t)
    side.synchronize()
    return ret if ret is not None else out.numpy()


def bind_host_to_device(device: Optional[Any] = None) -> Optional[str]:
    """Pin this process to the CPU cores next to ``device`` (its PCIe root's ``local_cpulist``),
    so that pinned result buffers are first-touched on the GPU's own NUMA node.  With one
    process per GPU on a two-socket box this keeps every device->host stream off the
    inter-socket link.  Returns the CPU list applied, or ``None`` when sysfs does not say."""
    import os

    try:
        d = default_device(device)
        pr = torch.cuda.get_device_properties(d)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        text = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return text
    except Exception:
        return None


def limits_tensor(limits: Any) -> Optional[torch.Tensor]:
    """(n, 2) float32 host tensor, or None when every bound is infinite (clip is a no-op)."""
    if limits is None:
        return None
    a = np.asarray(limits, dtype=np.float32)
    if a.size == 0 or (np.isneginf(a[:, 0]).all() and np.isposinf(a[:, 1]).all()):
        return None
    return torch.from_numpy(np.ascontiguousarray(a))


def vec(x: Any, n: int, name: str) -> list:
    a = np.asarray(x.detach().cpu() if isinstance(x, torch.Tensor) else x, dtype=np.float64).reshape(-1)
    if a.size != n:
        raise ValueError(f"{name} must have {n} entries, got {a.size}")
    return [float(v) for v in a]


def gravity(g: Any) -> list:
    return [0.0, 0.0, -9.81] if g is None else vec(g, 3, "gravity vector")


def is_f32(precision: Any) -> bool:
    """``precision`` keyword of the API mirrors: float64 (default: the reference's NumPy path,
    1e-9) or float32 arithmetic (the north-star's fp32 kernels: 1e-4 on torques, 1e-5 on poses)."""
    if precision is None:
        return False
    p = np.dtype(precision) if not isinstance(precision, str) else np.dtype(
        {"float64": np.float64, "fp64": np.float64, "f64": np.float64, "double": np.float64,
         "float32": np.float32, "fp32": np.float32, "f32": np.float32, "single": np.float32}.get(precision.lower(), precision))
    if p == np.float64:
        return False
    if p == np.float32:
        return True
    raise ValueError(f"precision must be float64 or float32, got {precision!r}")
