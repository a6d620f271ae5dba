#!/usr/bin/env python3
"""Static instruction count of the steady-state wavefront step of a strip kernel (development aid).

Finds, in the SASS of one kernel of libb200align.so, the innermost loops that hold the unrolled steady step
(recognised by their VIADDMNMX count), walks the fall-through path (the rare candidate-ring blocks guarded by a
forward branch are skipped) and prints instructions per step by opcode class -> instructions per DP cell.

  python tools/sass_step_count.py [--lib path] [--kernel mangled-substring] [--rows-per-lane 16]
"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ALU = ("VIADDMNMX", "VIMNMX3", "VIMNMX", "PRMT", "LOP3", "ISETP", "SEL", "SHF", "PLOP3", "LEA", "IADD3", "POPC", "FLO", "BREV", "IABS")
FMA = ("IMAD", "VIADD", "MOV", "FFMA")
MIO = ("LDS", "STS", "SHFL", "LDG", "STG", "LD", "ST", "ATOM", "RED", "S2R", "VOTE")


def klass(op):
    base = op.split(".")[0]
    if base.startswith("U") and base not in ("UNPACK",):
        return "uniform"
    if base in ("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "WARPSYNC", "NOP", "BAR"):
        return "branch"
    if base in MIO or base.startswith("LDS") or base.startswith("STS"):
        return "mio"
    if base in ("IMAD", "MOV", "FFMA") or (base == "VIADD"):
        return "fma"
    return "alu"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "masa-cudalign_b200", "libb200align.so"))
    ap.add_argument("--kernel", default="strip_kernel_s16ILi16ELb1ELb1ELb0")
    ap.add_argument("--rows-per-lane", type=int, default=16)
    ap.add_argument("--dump", action="store_true")
    args = ap.parse_args()
    out = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True, check=True).stdout
    # split per function
    funcs = re.split(r"\n\s+Function : ", out)
    body = next((f for f in funcs if args.kernel in f.split("\n", 1)[0]), None)
    if body is None:
        raise SystemExit(f"kernel {args.kernel} not found")
    ins = []
    for l in body.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    idx = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        if "BRA" in t and "BRA.DIV" not in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a and int(m.group(1), 16) in idx:
                loops.append((idx[int(m.group(1), 16)], i))
    R = args.rows_per_lane
    for s, e in loops:
        nv = sum(1 for _, t in ins[s:e + 1] if "VIADDMNMX" in t)
        steps = nv // (3 * R)
        if nv < 3 * R or nv % (3 * R) not in (0, steps) or e - s > 4000:      # + 1 per step: the F-chain bound of the FILT variant
            continue
        # walk the fall-through path
        i, path = s, []
        while i <= e:
            a, t = ins[i]
            path.append((a, t))
            if "BRA" in t and "BRA.DIV" not in t and t.startswith("@"):
                m = re.search(r"0x([0-9a-f]+)", t)
                tgt = int(m.group(1), 16) if m else None
                if tgt and tgt > a and tgt in idx and idx[tgt] <= e + 1:
                    blk = ins[i + 1:idx[tgt]]
                    if any("CALL" in x for _, x in blk) or sum("STS" in x for _, x in blk) >= 3:
                        i = idx[tgt]
                        continue
            i += 1
        c = collections.Counter()
        ops = collections.Counter()
        for a, t in path:
            t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
            op = t2.split()[0]
            c[klass(op)] += 1
            ops[op.split(".")[0] + (".16x2" if "16x2" in op.upper() else "")] += 1
        n = len(path)
        cells = 2 * R * steps
        print(f"loop 0x{ins[s][0]:x}-0x{ins[e][0]:x}: {steps} step(s), {e - s + 1} static, {n} on the common path = {n / steps:.1f}/step = "
              f"{n / cells:.3f} instr/cell; classes/step: " + ", ".join(f"{k} {v / steps:.1f}" for k, v in sorted(c.items())))
        print("   " + ", ".join(f"{k} {v / steps:.1f}" for k, v in ops.most_common(24)))
        if args.dump:
            for a, t in path:
                print(f"      {a:05x} {t}")


if __name__ == "__main__":
    main()
