#!/usr/bin/env python3
"""Instruction census of one kernel's SASS (no GPU needed: reads the object files nvcc wrote).

    python scripts/sass_census.py traj_rnea_kernel.*Li6E.*Lb0ELb1ELb0ELb0E [--obj dyn_flavour0.o] [--json]

Counts, for the function(s) whose mangled name matches the regular expression:
  * every mnemonic, grouped by the pipe it issues to (fp64, XU conversions / MUFU, LSU, integer / move, branch);
  * the fp64 instructions (DFMA / DMUL / DADD) by OPERAND PATTERN -- the quantity the fp64 pipe's
    sustained issue rate depends on (profiles/r1_variants.md J, `mpk_fma_peak` modes):
        uniform   one operand from a uniform register or the constant bank   (R, UR|c[..], R)
        imm       one immediate operand
        reuse     all vector registers, at least one flagged .reuse
        regs3     DFMA with three distinct vector-register operands, no reuse
        regs2     DMUL / DADD with two vector-register operands, no reuse
This is a STATIC count over the whole function body (straight-line code for the unrolled link
loops; slow paths such as the out-of-range sincos call are included, so the dynamic counts ncu
reports are a little lower).
"""

from __future__ import annotations

import argparse
import json
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

OBJ = Path(__file__).resolve().parents[1] / "manipulapy_b200" / "_lib" / "obj"

PIPES = [
    ("fp64", re.compile(r"^(DFMA|DMUL|DADD|DSETP|DMNMX)")),
    ("xu", re.compile(r"^(F2F|I2F|F2I|MUFU|FRND)")),
    ("lsu_shared", re.compile(r"^(LDS|STS|LDSM)")),
    ("lsu_global", re.compile(r"^(LDG|STG|LD\b|ST\b|LDL|STL|LDGSTS|LDGDEPBAR|RED|ATOM)")),
    ("const", re.compile(r"^(LDC|ULDC)")),
    ("branch", re.compile(r"^(BRA|BSSY|BSYNC|CALL|RET|EXIT|WARPSYNC|BAR|NANOSLEEP|DEPBAR|ERRBAR|MEMBAR)")),
    ("uniform", re.compile(r"^(U[A-Z0-9]+|R2UR|S2UR|VOTEU)")),
    ("fp32", re.compile(r"^(FFMA|FMUL|FADD|FSETP|FMNMX|FSEL)")),
    ("int_move", re.compile(r".*")),
]


def functions(obj: Path):
    text = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    name, body = None, []
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.+?);", line)
        if m and name:
            body.append(m.group(1) + " " + m.group(2).strip())
    if name:
        yield name, body


def rf_rows(op: str, operands: list) -> tuple:
    """(register-file rows read, rows flagged .reuse) of one instruction: one row = one 32-bit
    register x 32 lanes.  Measured on B200 (profiles/r2_fp64_operand_patterns.md): an SM sub-partition
    reads about two rows per cycle for ALL of its instructions, so sum(rows) / 2 is a cycle floor
    next to the fp64 pipe's 2 cycles per instruction."""
    base = op.split(".")[0]
    store = base in ("STS", "STG", "ST", "STL", "RED", "ATOM", "ATOMS", "ATOMG")
    srcs = operands if store else operands[1:]
    if base in ("ISETP", "FSETP", "DSETP", "PLOP3", "UISETP"):
        srcs = operands  # predicate destinations carry no R
    wide64 = base in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX") or ".F32.F64" in op or (
        base in ("F2I", "FRND", "I2F") and (".F64" in op and not op.endswith(".F64.F32")) and "S64" in op) or (
        base == "F2I" and ".F64" in op)
    rows = reuse = 0
    for k, a in enumerate(srcs):
        regs = re.findall(r"(?<![UP])R(\d+)", a)
        if not regs:
            continue
        w = 2 if wide64 else 1
        if ".64" in a:
            w = 2
        if store and k > 0:  # data operand of a store
            w = 4 if ".128" in op else 2 if ".64" in op else 1
        if base == "IMAD" and ".WIDE" in op and k == len(srcs) - 1:
            w = 2
        rows += w * len(regs)
        if ".reuse" in a:
            reuse += w
    return rows, reuse


def census(body, lo=None, hi=None):
    pipes, mnem, pat = Counter(), Counter(), Counter()
    rows_by_pipe, rows, reuse_rows = Counter(), 0, 0
    for ins in body:
        m = re.match(r"^([0-9a-f]{4,})\s+(.*)$", ins)
        if m:
            addr, ins = int(m.group(1), 16), m.group(2)
            if (lo is not None and addr < lo) or (hi is not None and addr >= hi):
                continue
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
        op = ins.split()[0]
        base = op.split(".")[0]
        mnem[base] += 1
        for pipe, rx in PIPES:
            if rx.match(op):
                pipes[pipe] += 1
                break
        r, ru = rf_rows(op, [a.strip() for a in ins[len(op):].split(",")])
        rows += r
        reuse_rows += ru
        rows_by_pipe[pipe] += r
        # per-instruction operand collection: ceil(rows / 2) cycles, at least the issue slot; an fp64
        # instruction holds its pipe for 2 cycles
        cyc = max(1, -(-r // 2))
        if pipe == "fp64":
            cyc = max(2, cyc)
        pat["_cycles_ceil_model"] += cyc
        if base in ("DFMA", "DMUL", "DADD"):
            args = ins[len(op):]
            srcs = [a.strip() for a in args.split(",")][1:]
            if any(re.search(r"\bUR\d+|c\[", a) for a in srcs):
                k = "uniform"
            elif any(re.fullmatch(r"-?\|?[-+0-9.e]+(?:e[-+]?\d+)?\|?|-?(?:\+)?INF|-?QNAN", a.replace(" ", "")) for a in srcs):
                k = "imm"
            elif any(".reuse" in a for a in srcs):
                k = "reuse"
            else:
                k = "regs3" if base == "DFMA" else "regs2"
            pat[f"{base}:{k}"] += 1
            pat[k] += 1
    n = sum(pipes.values())
    return {"instructions": n, "pipes": dict(pipes), "fp64_patterns": dict(pat),
            "rf_rows": rows, "rf_rows_reuse_flagged": reuse_rows, "rf_rows_by_pipe": dict(rows_by_pipe),
            "mnemonics": dict(mnem.most_common())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("regex")
    ap.add_argument("--obj", default=None, help="object file under manipulapy_b200/_lib/obj (default: all)")
    ap.add_argument("--json", action="store_true")
    ap.add_argument("--range", default=None, help="hex address range lo:hi of the function body to count")
    args = ap.parse_args()
    rx = re.compile(args.regex)
    objs = [OBJ / args.obj] if args.obj else sorted(OBJ.glob("*.o"))
    lo = hi = None
    if args.range:
        a, b = args.range.split(":")
        lo, hi = (int(a, 16) if a else None), (int(b, 16) if b else None)
    out = {}
    for o in objs:
        for name, body in functions(o):
            if rx.search(name):
                out[f"{o.name}:{name}"] = census(body, lo, hi)
    if args.json:
        print(json.dumps(out, indent=1))
        return
    for k, c in out.items():
        f = c["fp64_patterns"]
        n64 = sum(f.get(x, 0) for x in ("uniform", "imm", "reuse", "regs3", "regs2"))
        print(k)
        print(f"  instructions {c['instructions']}: " + ", ".join(f"{p} {v}" for p, v in sorted(c['pipes'].items(), key=lambda kv: -kv[1])))
        if n64:
            print(f"  fp64 {n64}: " + ", ".join(
                f"{x} {f.get(x, 0)} ({100 * f.get(x, 0) / n64:.0f} %)" for x in ("uniform", "imm", "reuse", "regs3", "regs2")))
            print("    DFMA {} DMUL {} DADD {}".format(*(c["mnemonics"].get(m, 0) for m in ("DFMA", "DMUL", "DADD"))))
        print(f"  register-file rows read {c['rf_rows']} (of which flagged .reuse {c['rf_rows_reuse_flagged']}): "
              + ", ".join(f"{p} {v}" for p, v in sorted(c['rf_rows_by_pipe'].items(), key=lambda kv: -kv[1]) if v)
              + f"  => cycle floors per warp: rows/2 = {(c['rf_rows'] - c['rf_rows_reuse_flagged']) / 2:.0f}"
              + (f", fp64 pipe 2 x {n64} = {2 * n64}" if n64 else "")
              + f", sum of per-instruction max(issue, ceil(rows/2)) = {f.get('_cycles_ceil_model', 0)}")
    if not out:
        sys.exit("no function matched")


if __name__ == "__main__":
    main()
