#!/usr/bin/env python3
"""fp64 issue rate of a LONE warp (and of 2 .. 16 warps per SM): cycles per DFMA warp instruction.
The forward-dynamics rollouts at small batch sizes run with about one warp per scheduler; whether such a
warp can feed its scheduler's 16-lane fp64 unit back to back (one instruction per 2 cycles) or only every
other slot decides what a finer split of the step across warps can gain.

    python scripts/lone_warp_probe.py > gpurun_out/lone_warp.json
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from manipulapy_b200 import _native  # noqa: E402


def main() -> None:
    ops = _native.ops()
    dev = torch.device("cuda:0")
    sink = torch.zeros(1, dtype=torch.float64, device=dev)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    clock_hz = 1.965e9
    out = {"sms": sms, "assumed_clock_hz": clock_hz, "rows": []}
    iters = 1 << 16
    for mode, label, instr in ((0, "8 independent chains, shared operands", 8), (1, "three distinct registers (1 chain of 8)", 8),
                               (2, "constant-bank operand", 8)):
        for threads in (32, 64, 128, 256, 512):
            for blocks_per_sm in (1, 2):
                blocks = sms * blocks_per_sm
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                best = 1e30
                for _ in range(4):
                    ev0.record()
                    ops.fma_peak(sink, mode << 8, blocks, threads, iters)
                    ev1.record()
                    torch.cuda.synchronize()
                    best = min(best, ev0.elapsed_time(ev1) * 1e-3)
                warps_per_sm = threads // 32 * blocks_per_sm
                cycles_per_instr_per_warp = best * clock_hz / (iters * instr)
                out["rows"].append({"pattern": label, "threads": threads, "blocks_per_sm": blocks_per_sm,
                                    "warps_per_sm": warps_per_sm, "ms": best * 1e3,
                                    "cycles_per_dfma_per_warp": cycles_per_instr_per_warp,
                                    "dfma_per_cycle_per_sm": warps_per_sm / cycles_per_instr_per_warp})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
