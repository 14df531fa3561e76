#!/usr/bin/env python3
"""Multi-GPU functional check (launch with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 scripts/multigpu_check.py

Every rank computes its contiguous shard of a trajectory + inverse-dynamics batch and of a
forward-dynamics rollout batch; the shards are gathered over NCCL (all-gather and gather-to-0)
and compared bit for bit with the whole batch computed on rank 0's GPU alone.  Prints one JSON
line on rank 0 with the gather bandwidth."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from manipulapy_b200 import _native, gather_rows, load_robot, shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ops = _native.ops()
    ur5 = load_robot("ur5", device=dev)
    planner = ur5.planner()
    rng = np.random.default_rng(5)
    B, N = 1003, 257  # ragged: not divisible by the world size
    s_all, e_all = rng.uniform(-3, 3, (B, 6)), rng.uniform(-3, 3, (B, 6))
    lo, hi = shard_range(B, world, rank)
    s, e = torch.from_numpy(s_all[lo:hi]).to(dev), torch.from_numpy(e_all[lo:hi]).to(dev)
    tau = planner.trajectory_inverse_dynamics(s, e, 2.0, N, 5)
    everyone = gather_rows(tau, B)                      # all-gather over NVLink
    on0 = gather_rows(tau, B, dst=0)                    # gather to rank 0
    ok = True
    if rank == 0:
        whole = planner.trajectory_inverse_dynamics(torch.from_numpy(s_all).to(dev), torch.from_numpy(e_all).to(dev),
                                                    2.0, N, 5)
        ok &= bool(torch.equal(whole.view(torch.int32), on0.view(torch.int32)))
        ok &= bool(torch.equal(whole.view(torch.int32), everyone.view(torch.int32)))
    else:
        ok &= on0 is None and everyone.shape == (B, N, 6)
    # rollouts
    iiwa = load_robot("iiwa14", device=dev)
    Bf, Nf = 20011, 24  # whole batch: one warp per 32 rollouts; shards: warp-pair kernel (same bits)
    th0 = rng.uniform(-1, 1, (Bf, 7)); dth0 = rng.uniform(-0.5, 0.5, (Bf, 7)); tm = rng.uniform(-5, 5, (Bf, Nf, 7)).astype(np.float32)
    lo, hi = shard_range(Bf, world, rank)
    pl7 = iiwa.planner()
    r = pl7.forward_dynamics_trajectory(torch.from_numpy(th0[lo:hi]).to(dev), torch.from_numpy(dth0[lo:hi]).to(dev),
                                        torch.from_numpy(tm[lo:hi]).to(dev), [0, 0, -9.81], None, 1e-3, 1)
    pos = gather_rows(r["positions"], Bf, dst=0)
    if rank == 0:
        w = pl7.forward_dynamics_trajectory(torch.from_numpy(th0).to(dev), torch.from_numpy(dth0).to(dev),
                                            torch.from_numpy(tm).to(dev), [0, 0, -9.81], None, 1e-3, 1)
        ok &= bool(torch.equal(w["positions"].view(torch.int32), pos.view(torch.int32)))
    # gather bandwidth on a result-sized buffer (4096 x 2441 x 6 float32 per rank)
    big = torch.empty((4096 // world * world // world, 2441, 6), dtype=torch.float32, device=dev).normal_()
    gather_rows(big, big.shape[0] * world)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(5):
        gather_rows(big, big.shape[0] * world)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "bit_identical_to_single_gpu": bool(flag.item()),
                          "all_gather_ms": dt * 1e3,
                          "all_gather_gbs_per_rank_out": big.numel() * 4 * (world - 1) / dt / 1e9}))
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
