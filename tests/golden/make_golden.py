#!/usr/bin/env python3
"""Regenerates tests/golden/*.json from the REFERENCE's own code: oracle/_ref/oracle_cpu (MASA-Core 1.4.2.1028 +
CPUBlockProcessor under the AbstractDiagonalAligner policy, compiled from /root/reference by oracle/Makefile).
The reference ships no golden vectors, so these fixtures are what pins the C restatement
(oracle/gotoh_oracle.c) and the CUDA path.  Run in the build container (needs oracle/_ref):

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import struct
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
import oracle_lib as O  # noqa: E402

CASES = {
    # name: (m, n, (a0, a1), ps, pd, pi, seed, extra flags)
    "sw_3k": (3000, 2700, (500, 2500), 0.05, 0.02, 0.02, 11, []),
    "sw_12k_rows": (12000, 9000, (1000, 11000), 0.05, 0.02, 0.02, 12, ["--no-block-pruning", "--disk-size=1M"]),
    "sw_40k": (40000, 39979, (5000, 35000), 0.05, 0.02, 0.02, 11, ["--no-block-pruning", "--disk-size=4M"]),
    "nw_20k": (20000, 21000, (0, 20000), 0.05, 0.02, 0.02, 13, ["--alignment-edges=++", "--no-block-pruning", "--disk-size=4M"]),
}


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    out = {}
    for name, (m, n, hom, ps, pd, pi, seed, extra) in CASES.items():
        a, b = synth.make_pair(m, n, [hom], ps, pd, pi, 0, seed)
        with tempfile.TemporaryDirectory() as td:
            fa, fb = os.path.join(td, "A.fa"), os.path.join(td, "B.fa")
            synth.write_fasta(fa, a, "A"); synth.write_fasta(fb, b, "B")
            wd = os.path.join(td, "w")
            O.run_ref("oracle_cpu", fa, fb, wd, extra)
            xp = {}
            for f in sorted(os.listdir(os.path.join(wd, "crosspoints"))):
                xp[f] = O.read_crosspoints(os.path.join(wd, "crosspoints", f))
            rows = {}
            srd = os.path.join(wd, "special_rows", "stage.01.00")
            if os.path.isdir(srd):
                for base, _d, names in os.walk(srd):
                    for fn in sorted(names):
                        p = os.path.join(base, fn)
                        raw = open(p, "rb").read()
                        if len(raw) < 8:
                            continue
                        first = struct.unpack("<ii", raw[:8])
                        rows[os.path.relpath(p, srd)] = {"sha256": hashlib.sha256(raw).hexdigest(), "cells": len(raw) // 8, "first_cell": list(first)}
            out[name] = {
                "generator": {"m": m, "n": n, "homology": list(hom), "p_s": ps, "p_d": pd, "p_i": pi, "seed": seed},
                "flags": extra,
                "seq_sha256": [hashlib.sha256(a.tobytes()).hexdigest(), hashlib.sha256(b.tobytes()).hexdigest()],
                "crosspoints": xp,
                "alignment_bin_sha256": sha(os.path.join(wd, "alignment.00.bin")),
                "special_rows_stage1": rows,
            }
            print(name, "stage1:", xp.get("crosspoint_01.00"), "rows:", len(rows))
    with open(os.path.join(HERE, "reference_runs.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
