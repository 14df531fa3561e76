"""The fused gather (``sharding.PeerRows``: every rank's kernel stores its result rows straight into the
collecting rank's buffer, mapped through CUDA IPC) across TWO PROCESSES.  The GPU test box has one GPU, so
both processes use ``cuda:0`` -- the IPC export / import, the aligned shard bounds, the store into a mapped
buffer of another process and the ``commit()`` ordering are the same code that runs over NVLink on a
multi-GPU box (there: ``bench.py``'s cfg 5 sweep, NCCL for the control plane); the control plane here is
gloo, because NCCL refuses two ranks on one device.  Bar: the gathered rows equal a single-process
computation of the whole batch bit for bit, whichever rank collects.  (The NCCL pipeline is covered on the CPU
under gloo -- tests/test_native_cpu.py -- and measured by bench.py on 2 / 4 / 8 GPUs; gloo has no device send / recv.)"""

import os
import subprocess
import sys
from pathlib import Path

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

REPO = Path(__file__).resolve().parents[1]

_WORKER = r"""
import sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from manipulapy_b200 import _native, load_robot
from manipulapy_b200.sharding import PeerRows, shard_bounds
rank = int(sys.argv[3])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=2)
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
ops = _native.ops()
rb = load_robot("ur5", device=dev)
handle, jl = rb.dynamics.robot.handle, rb.planner()._jl
B, N, TF, METHOD, g = 37, 61, 2.0, 5, [0.0, 0.0, -9.81]        # ragged: 2257 points, odd shard sizes
gen = torch.Generator(device=dev).manual_seed(5)
ends = (torch.rand(2, B, 6, dtype=torch.float64, device=dev, generator=gen) * 2 - 1) * np.pi

def launch(lo, hi, dest):
    if hi > lo:
        ops.trajectory_inverse_dynamics(handle, ends[0, lo:hi].contiguous(), ends[1, lo:hi].contiguous(), False, TF, N,
                                        METHOD, jl, g, None, None, False, False, dest)

pr = PeerRows(B, (N, 6), torch.float32, dev, dst=0)
assert pr.bounds == shard_bounds(B, 2, None, pr.align) and pr.align == 2   # 61 x 6 x 4 bytes per trajectory = 8 mod 16
if rank == 0:
    pr.full.fill_(float("nan"))
dist.barrier()
launch(pr.lo, pr.hi, pr.rows())
pr.commit()
torch.cuda.synchronize()
dist.barrier()
ref = torch.empty((B, N, 6), dtype=torch.float32, device=dev)
launch(0, B, ref)
if rank == 0:
    assert torch.equal(pr.full.view(torch.int32), ref.view(torch.int32)), "peer-stored rows differ from the single-process result"
# a second buffer with the shards the other way round (weights): rank 1 owns the larger, first-unaligned part
pr2 = PeerRows(B, (N, 6), torch.float32, dev, dst=1, weights=[1.0, 3.0])
launch(pr2.lo, pr2.hi, pr2.rows())
pr2.commit()
torch.cuda.synchronize()
dist.barrier()
if rank == 1:
    assert pr2.lo % 2 == 0 and torch.equal(pr2.full.view(torch.int32), ref.view(torch.int32))
pr2.close()
pr.close()
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_peer_rows_two_processes_one_gpu(tmp_path):
    script = tmp_path / "peer_worker.py"
    script.write_text(_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(REPO), port, str(r)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in out, out[-3000:]
