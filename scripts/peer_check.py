#!/usr/bin/env python3
"""Which way of mapping another rank's HBM works on this box?  Run under torchrun with >= 2 ranks:

    timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 scripts/peer_check.py

Every rank prints one JSON line: a kernel of this package storing into rank 0's buffer through the
CUDA IPC mapping of `PeerRows` (bits compared with a single-GPU run), and torch symmetric memory.
"""

import datetime
import json
import os
import sys
import traceback
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=60))
    res = {"rank": rank, "peer_access_to_0": torch.cuda.can_device_access_peer(local, 0) if local else None}

    # 2. this package's kernel storing through the mapping (PeerRows)
    try:
        from manipulapy_b200 import _native, load_robot
        from manipulapy_b200.sharding import PeerRows

        ops = _native.ops()
        rb = load_robot("ur5", device=dev)
        B, N = 64, 257
        pr = PeerRows(B, (N, 6), torch.float32, dev, dst=0)
        gen = torch.Generator(device=dev).manual_seed(1)
        ends = torch.rand(2, B, 6, dtype=torch.float64, device=dev, generator=gen) * 2 - 1
        jl = rb.planner()._jl
        ops.trajectory_inverse_dynamics(rb.dynamics.robot.handle, ends[0, pr.lo:pr.hi].contiguous(),
                                        ends[1, pr.lo:pr.hi].contiguous(), False, 2.0, N, 5, jl, [0.0, 0.0, -9.81], None, None,
                                        False, False, pr.rows())
        pr.commit()
        torch.cuda.synchronize()
        if rank == 0:
            ref = ops.trajectory_inverse_dynamics(rb.dynamics.robot.handle, ends[0], ends[1], False, 2.0, N, 5, jl,
                                                  [0.0, 0.0, -9.81], None, None, False)[0]
            res["peer_rows_bits_equal_single_gpu"] = bool(torch.equal(ref, pr.full))
        pr.close()
        res["peer_rows"] = "ok"
    except Exception as ex:
        res["peer_rows"] = f"{type(ex).__name__}: {ex}"
        res["peer_rows_trace"] = traceback.format_exc()[-600:]

    # 3. torch symmetric memory
    try:
        import torch.distributed._symmetric_memory as symm

        t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD)
        buf0 = hdl.get_buffer(0, (1 << 20,), torch.float32)
        buf0[rank << 18:(rank + 1) << 18].fill_(float(rank) + 10)
        hdl.barrier()
        torch.cuda.synchronize()
        if rank == 0:
            res["symm_values_seen_on_0"] = [float(t[r << 18]) for r in range(min(world, 4))]
        res["symm"] = "ok"
    except Exception as ex:
        res["symm"] = f"{type(ex).__name__}: {ex}"[:400]
    print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
