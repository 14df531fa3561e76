#!/usr/bin/env python3
"""Concurrent device->host bandwidth of the box: every rank copies a 256 MiB device buffer into
pinned host memory at the same time (barrier first).  Explains the end-to-end number of
``bench.py --gpus N``, whose result rows (240 MB per step and GPU) all cross PCIe.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/d2h_probe.py
"""

import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 256 << 20
    src = torch.empty(n, dtype=torch.uint8, device=dev)
    pin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    pin.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    res = {}
    for mode in ("alone", "together"):
        rates = []
        for r in range(world if mode == "alone" else 1):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            if mode == "together" or r == rank:
                t0 = time.perf_counter()
                for _ in range(8):
                    pin.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
                rates.append(8 * n / (time.perf_counter() - t0) / 1e9)
        t = torch.tensor([rates[0]], dtype=torch.float64, device=dev)
        if world > 1:
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            res[mode] = [round(float(x), 2) for x in out]
        else:
            res[mode] = [round(float(t), 2)]
    if rank == 0:
        res["sum_together_gbs"] = round(sum(res["together"]), 1)
        res["world"] = world
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
