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
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.+?);", line)
        if m and name:
            body.append(m.group(1).strip())
    if name:
        yield name, body


def census(body):
    pipes, mnem, pat = Counter(), Counter(), Counter()
    for ins in body:
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
        op = ins.split()[0]
        base = op.split(".")[0]
        mnem[base] += 1
        for pipe, rx in PIPES:
            if rx.match(op):
                pipes[pipe] += 1
                break
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
    return {"instructions": len(body), "pipes": dict(pipes), "fp64_patterns": dict(pat),
            "mnemonics": dict(mnem.most_common())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("regex")
    ap.add_argument("--obj", default=None, help="object file under manipulapy_b200/_lib/obj (default: all)")
    ap.add_argument("--json", action="store_true")
    args = ap.parse_args()
    rx = re.compile(args.regex)
    objs = [OBJ / args.obj] if args.obj else sorted(OBJ.glob("*.o"))
    out = {}
    for o in objs:
        for name, body in functions(o):
            if rx.search(name):
                out[f"{o.name}:{name}"] = census(body)
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
    if not out:
        sys.exit("no function matched")


if __name__ == "__main__":
    main()
