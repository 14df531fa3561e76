"""fp64 issue rate of a lone warp with 32 / 16 / 8 active lanes (does the 16-lane pipe skip an empty half warp?  No.)"""
import sys, json, torch
sys.path.insert(0, '.')
from manipulapy_b200 import _native
ops = _native.ops(); dev = torch.device('cuda:0')
sink = torch.zeros(1, dtype=torch.float64, device=dev)
sms = torch.cuda.get_device_properties(0).multi_processor_count
iters = 1 << 16
for mode in (0, 1):
    for threads in (32, 16, 8):
        best = 1e9
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.fma_peak(sink, mode << 8, sms, threads, iters); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) * 1e-3)
        print(json.dumps({"mode": mode, "active_lanes": threads, "cycles_per_dfma": best * 1.965e9 / (iters * 8)}))
