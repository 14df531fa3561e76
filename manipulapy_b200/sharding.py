"""Multi-GPU sharding of the hot path.

Every unit of work (a trajectory point, a configuration, a whole rollout) is independent,
so the batch is split into contiguous index ranges, one per rank (one process per GPU,
``torch.distributed``); robot constants are replicated (< 4 KB) and each rank computes its
own slice.  There is no reduction and no exchange step; the only data movement between GPUs
is the optional final gather of the result rows onto one rank.  Three ways to do it:

``PeerRows``               the collecting rank's result buffer is mapped into every other rank's
                           address space (CUDA IPC over NVLink / NVSwitch); each rank's kernel
                           stores its rows straight into it, so the gather is FUSED into the
                           kernel -- every coalesced tile goes over the link as it is produced
                           and there is no separate transfer step.
``gather_rows_pipelined``  NCCL baseline: the shard is computed in chunks and every finished
                           chunk is sent (``isend`` / ``irecv`` on a second stream) while the next
                           one is being computed.
``gather_rows``            NCCL all-gather / gather of the finished shard (one shot, no overlap).
"""

from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(units: int, world_size: int, weights: Optional[Sequence[float]] = None, align: int = 1) -> list:
    """``world_size + 1`` row offsets of the contiguous shards.  ``weights=None``:
    ``ceil(units / world)`` rows per rank.  With weights (e.g. each rank's measured device->host
    rate when the results are host-destined) rank r gets a share proportional to ``weights[r]``.
    ``align``: interior boundaries are multiples of ``align`` units (so that every shard of a shared
    result buffer starts on a 16-byte boundary and is written with full-width vector stores)."""
    units, world_size, align = int(units), int(world_size), max(1, int(align))
    if world_size < 1:
        raise ValueError("bad world_size")
    if weights is None:
        per = -(-units // world_size)
        per = -(-per // align) * align
        return [min(units, r * per) for r in range(world_size)] + [units]
    w = [float(x) for x in weights]
    if len(w) != world_size or any(not (x > 0) for x in w):
        raise ValueError("weights must be world_size positive numbers")
    total, acc, out = sum(w), 0.0, [0]
    for r in range(world_size - 1):
        acc += w[r]
        b = int(round(units * acc / total / align)) * align
        out.append(min(units, max(out[-1], b)))
    return out + [units]


def shard_range(units: int, world_size: int, rank: int,
                weights: Optional[Sequence[float]] = None, align: int = 1) -> Tuple[int, int]:
    """Contiguous ``[lo, hi)`` of ``units`` owned by ``rank`` (see ``shard_bounds``)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    b = shard_bounds(units, world_size, weights, align)
    return b[rank], b[rank + 1]


def _world(group) -> Tuple[int, int]:
    if not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def gather_rows(local: torch.Tensor, units: int, group: Optional[dist.ProcessGroup] = None,
                dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """Concatenate the per-rank row slices produced under ``shard_range`` back into ``units`` rows.

    ``dst=None``: all-gather (every rank gets the result); otherwise only ``dst`` does
    (``None`` elsewhere).  Slices are padded to the common ``ceil(units / world)`` rows so a
    single fixed-size collective moves everything.
    """
    world, rank = _world(group)
    if world == 1:
        return local
    per = -(-int(units) // world)
    tail = local.shape[1:]
    buf = local
    if local.shape[0] != per:
        buf = local.new_zeros((per, *tail))
        buf[: local.shape[0]] = local
    buf = buf.contiguous()
    if dst is None:
        out = local.new_empty((world * per, *tail))
        dist.all_gather_into_tensor(out, buf, group=group)
        return out[:units]
    pieces = [local.new_empty((per, *tail)) for _ in range(world)] if rank == dst else None
    dist.gather(buf, pieces, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat(pieces, 0)[:units]


def gather_rows_pipelined(launch: Callable[[int, int, torch.Tensor], None], units: int, tail: Sequence[int],
                          dtype: torch.dtype, device: torch.device, dst: int = 0, chunks: int = 8,
                          group: Optional[dist.ProcessGroup] = None, out: Optional[torch.Tensor] = None,
                          side: Optional[torch.cuda.Stream] = None, align: int = 1) -> Optional[torch.Tensor]:
    """Compute this rank's shard of ``units`` rows in ``chunks`` pieces and gather them on ``dst``
    while computing: ``launch(lo, hi, dest)`` enqueues the kernel for global rows ``[lo, hi)`` writing
    into ``dest`` (``(hi - lo, *tail)``) on the current stream; every finished chunk is handed to NCCL
    (``isend`` on the producers, ``irecv`` straight into the final buffer on ``dst``) on a second
    stream, so the transfer of chunk k overlaps the kernel of chunk k + 1.  Returns the full
    ``(units, *tail)`` tensor on ``dst`` (``out`` if given), ``None`` elsewhere.  ``align``: shard
    boundaries as ``shard_bounds(..., align=align)`` gives them.  Works with any backend (the CPU
    tests run it over gloo)."""
    world, rank = _world(group)
    bounds = shard_bounds(units, world, None, align)
    lo, hi = bounds[rank], bounds[rank + 1]
    cuda = torch.device(device).type == "cuda"
    if rank == dst:
        full = out if out is not None else torch.empty((units, *tail), dtype=dtype, device=device)
    per = max(1, -(-max(b - a for a, b in zip(bounds, bounds[1:])) // max(1, chunks)))
    if cuda:
        compute = torch.cuda.current_stream(device)
        side = side if side is not None else torch.cuda.Stream(device=device)
    pending, keep = [], []
    nsteps = -(-max(b - a for a, b in zip(bounds, bounds[1:])) // per) if units else 0
    for k in range(nsteps):
        a, b = min(hi, lo + k * per), min(hi, lo + (k + 1) * per)
        if rank == dst:
            if b > a:
                launch(a, b, full[a:b])
        elif b > a:
            piece = torch.empty((b - a, *tail), dtype=dtype, device=device)
            launch(a, b, piece)
            keep.append(piece)
        if world == 1:
            continue
        if cuda:
            done = torch.cuda.Event()
            done.record(compute)
            side.wait_event(done)
        ops = []
        if rank == dst:
            for r in range(world):
                if r == dst:
                    continue
                ra, rb_ = min(bounds[r + 1], bounds[r] + k * per), min(bounds[r + 1], bounds[r] + (k + 1) * per)
                if rb_ > ra:
                    ops.append(dist.P2POp(dist.irecv, full[ra:rb_], r, group))
        elif b > a:
            ops.append(dist.P2POp(dist.isend, keep[-1], dst, group))
        if ops:
            if cuda:
                with torch.cuda.stream(side):
                    pending += dist.batch_isend_irecv(ops)
            else:
                pending += dist.batch_isend_irecv(ops)
    for w in pending:
        w.wait()
    if cuda and world > 1:
        compute.wait_stream(side)
        for t in keep:
            t.record_stream(side)
    return full if rank == dst else None


class PeerRows:
    """A ``(units, *tail)`` float32 result buffer that lives in ONE rank's HBM (``dst``) and that every
    rank's kernels write their own contiguous row range into directly.

    ``dst`` allocates it (``mpk_peer_alloc``: ``cudaMalloc`` + CUDA IPC export) and broadcasts the
    64-byte handle over the process group; every other rank maps it into its own address space
    (``mpk_peer_open``; the driver enables peer access over NVLink / NVSwitch) and gets a tensor of
    its OWN device whose bytes live on ``dst``.  A kernel launched with that tensor as its output
    pointer stores over the link: the fused trajectory + inverse-dynamics kernel flushes its rows
    with fully coalesced 16-byte stores, so every tile crosses as it is produced -- there is no
    staging copy and no separate collective; compute and "gather" are one kernel.  ``commit()`` (a
    one-element all-reduce enqueued behind the kernels on every rank) orders the remote stores
    before whatever ``dst`` enqueues next.

    Raises ``RuntimeError`` on EVERY rank when any rank cannot map the buffer (decided by an
    all-reduce, so the ranks never part ways); callers then fall back to ``gather_rows_pipelined``.
    """

    def __init__(self, units: int, tail: Sequence[int], dtype: torch.dtype, device: torch.device,
                 dst: int = 0, group: Optional[dist.ProcessGroup] = None,
                 weights: Optional[Sequence[float]] = None):
        from . import _native

        if dtype != torch.float32:
            raise ValueError("PeerRows holds float32 rows")
        ops = _native.ops()
        self.world, self.rank = _world(group)
        self.group, self.dst, self.units, self.tail = group, dst, int(units), tuple(int(x) for x in tail)
        self.device = torch.device(device)
        # every rank's slice starts on a 16-byte boundary: full-width vector stores over the link
        row_bytes = 4
        for x in self.tail:
            row_bytes *= x
        import math

        self.align = 16 // math.gcd(16, row_bytes)
        self.bounds = shard_bounds(units, self.world, weights, self.align)
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        shape = (self.units, *self.tail)
        numel = 1
        for x in shape:
            numel *= x
        like = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._token = like
        self.full = self.peer = None
        if self.world == 1:
            self.full = self.peer = torch.empty(shape, dtype=dtype, device=self.device)
            return
        err, handle = None, [None]
        if self.rank == dst:
            try:
                flat, h = ops.peer_alloc(max(numel, 1), like)
                self.full = self.peer = flat[:numel].view(shape)
                handle = [bytes(h.numpy().tobytes())]
            except Exception as ex:
                err = f"rank {self.rank}: {type(ex).__name__}: {ex}"
        dist.broadcast_object_list(handle, src=dst, group=group)
        if self.rank != dst and handle[0] is not None:
            try:
                h = torch.frombuffer(bytearray(handle[0]), dtype=torch.uint8)
                self.peer = ops.peer_open(h, max(numel, 1), like)[:numel].view(shape)
            except Exception as ex:  # decided collectively below: the ranks must not part ways
                err = f"rank {self.rank}: {type(ex).__name__}: {ex}"
        elif handle[0] is None:
            err = err or "the collecting rank could not export its buffer"
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok) == 0:
            self.full = self.peer = None
            raise RuntimeError(err or "another rank could not map the collecting rank's buffer")

    def rows(self, lo: Optional[int] = None, hi: Optional[int] = None) -> torch.Tensor:
        """Destination view of global rows ``[lo, hi)`` (default: this rank's shard) -- local memory
        on ``dst``, peer-mapped memory elsewhere."""
        lo = self.lo if lo is None else lo
        hi = self.hi if hi is None else hi
        return self.peer[lo:hi]

    def commit(self) -> None:
        """Stream-ordered on every rank behind its kernels: when it completes on ``dst`` every rank's
        rows have landed in ``full``."""
        if self.world > 1:
            dist.all_reduce(self._token, group=self.group)

    def close(self) -> None:
        """Unmap on the producers first, then free on ``dst``."""
        if self.world > 1:
            torch.cuda.synchronize(self.device)
            if self.rank != self.dst:
                self.peer = None
            dist.barrier(group=self.group)
        self.peer = self.full = None
