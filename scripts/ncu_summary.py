#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the GPU-less build box) into a compact markdown table.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--md profiles/x.md]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs(blocks)"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem(blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "pipe_fp64_%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe_xu_%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "pipe_tensor_%"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wavefronts_%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("sm__icc_request_hit_rate.pct", "icache_hit_%"),
    ("gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "gcc_instr_req_%"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible_warps/cycle"),
]
STALLS = "smsp__average_warps_issue_stalled_{}_per_issue_active.ratio"
STALL_NAMES = ["wait", "no_instruction", "not_selected", "math_pipe_throttle", "short_scoreboard",
               "long_scoreboard", "dispatch_stall", "barrier", "mio_throttle", "lg_throttle", "branch_resolving"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "").replace("mpk::", "")
        lines.append(f"### `{short}`  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
        lines.append("| metric | value |\n|---|---|")
        for key, label in KEYS:
            if key not in idx:
                continue
            v, u = r[idx[key]], units[idx[key]]
            if label.startswith("dram_") and label != "dram_%":
                lines.append(f"| {label} | {to_bytes(v, u) / 1e6:.3f} MB |")
            else:
                lines.append(f"| {label} | {v} {u} |")
        st = []
        for s in STALL_NAMES:
            k = STALLS.format(s)
            if k in idx:
                st.append((float(r[idx[k]]), s))
        st.sort(reverse=True)
        lines.append("| top stalls (warps per issue) | " + ", ".join(f"{s} {v:.2f}" for v, s in st[:6]) + " |")
        lines.append("")
    text = "\n".join(lines)
    if "--md" in sys.argv:
        open(sys.argv[sys.argv.index("--md") + 1], "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
