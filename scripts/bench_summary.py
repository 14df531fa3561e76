#!/usr/bin/env python3
"""One-screen summary of bench.py JSON lines:  python scripts/bench_summary.py out1.json [out2.json ...]"""
import json, sys
for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(f"== {path}: n_gpus {d['n_gpus']}  value {d['value']:.4e} {d['unit']}  {d['ms_per_step']:.4f} ms/step  "
          f"e2e {d['e2e']['value']:.4e}  clocks {d['clocks']['sm_mhz']} MHz {d['clocks']['reasons']}")
    e = d["e2e"]
    print(f"   e2e sharding: {str(e.get("sharding"))[-80:]}  d2h together {[round(x, 1) for x in e['pcie_d2h_gbs_all_ranks_together']]} "
          f"bound {e['pcie_bound_points_per_s']:.3e} (equal shards {e['pcie_bound_points_per_s_equal_shards']:.3e})")
    r = d["roofline"]
    print(f"   roofline hbm {r['frac']:.3f}  fp64 flop frac {r['fp64']['frac']:.3f}  issue frac {r['fp64']['pipe_issue_frac']:.3f}")
    if "fd_rollout" in d:
        f = d["fd_rollout"]
        print(f"   fd_rollout {f['value']:.3e} steps/s  {f['ms_per_launch']:.3f} ms  8192: {f['ms_per_launch_8192_rollouts']:.3f} ms")
    for c in d.get("configs", []):
        if "sizes" in c:
            for s in c["sizes"]:
                extra = ""
                if "peer_store_ms" in s:
                    extra = (f" peer {s['peer_store_ms']:.3f} ms nccl {s['nccl_pipelined_ms']:.3f} ms xfer {s['nccl_transfer_only_ms']:.3f} ms "
                             f"ovl {s['nccl_overlap_fraction']:.2f} bits {s.get('gathered_equals_single_gpu_bits')}")
                if "peer_store_error" in s:
                    extra += " PEER ERR " + s["peer_store_error"]
                print(f"   cfg5 P={s['points']:.3e} compute {s['compute_ms']:.3f} ms ({s['points_per_s_compute_only']:.3e}/s) "
                      f"gathered {s['gathered_points_per_s']:.3e}/s{extra} clk {s['clocks']['samples'] if s['clocks'] else None}")
        else:
            rf = c.get("roofline", {})
            print(f"   {c['name']}: {c['ms']:.4f} ms  {c['value']:.3e} {c['unit']}  hbm {rf.get('frac', 0):.3f} clk samples {c['clocks']['samples'] if c.get('clocks') else None}")
    if "cpu_baseline" in d:
        print("   cpu", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["cpu_baseline"].items() if k not in ("sample", "python_reference_source")})
