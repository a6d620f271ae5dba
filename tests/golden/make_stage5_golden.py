#!/usr/bin/env python3
"""Regenerates tests/golden/stage5_runs.json from the REFERENCE's own stage 5 (C/stage5/sw_stage5.cpp): runs
oracle/_ref/oracle_cpu (stages 1-6 of MASA-Core compiled from /root/reference) on small synthetic pairs and stores
  * the stage-4 crosspoints that stage 5 consumed (crosspoint_04.00),
  * what stage 5 produced, read back from alignment.00.bin with the reference's own reader
    (oracle/_ref/dump_alignment): raw score, match/mismatch/gap counters, start/end, both gap lists.
These vectors pin oracle/gotoh_oracle.c's go_stage5 (tests/test_oracle_cpu.py) and, through it and directly, the CUDA
traceback (tests/test_stage5_gpu.py).  Run in the build container (needs oracle/_ref):

    python tests/golden/make_stage5_golden.py
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
import oracle_lib as O  # noqa: E402

CASES = {
    # name: (m, n, [(a0, a1)], ps, pd, pi, K large indels, seed, extra flags)
    "sw_3k": (3000, 2700, [(500, 2500)], 0.05, 0.02, 0.02, 0, 11, []),
    "sw_6k_gappy": (6000, 6500, [(300, 5700)], 0.08, 0.06, 0.06, 0, 21, []),
    "sw_30k_indels_mps64": (30000, 30000, [(2000, 28000)], 0.04, 0.01, 0.01, 6, 22, ["--maximum-partition=64"]),
    "nw_5k_global": (5000, 5200, [(0, 5000)], 0.05, 0.02, 0.02, 0, 23, ["--alignment-edges=++"]),
    "sw_20k_mps700": (20000, 19000, [(1000, 19000)], 0.06, 0.02, 0.02, 2, 24, ["--maximum-partition=700"]),
}


def main():
    out = {}
    for name, (m, n, segs, ps, pd, pi, K, seed, extra) in CASES.items():
        a, b = synth.make_pair(m, n, segs, ps, pd, pi, K, seed)
        with tempfile.TemporaryDirectory() as td:
            fa, fb = os.path.join(td, "A.fa"), os.path.join(td, "B.fa")
            synth.write_fasta(fa, a, "A"); synth.write_fasta(fb, b, "B")
            wd = os.path.join(td, "w")
            O.run_ref("oracle_cpu", fa, fb, wd, extra)
            pts = O.read_crosspoints(os.path.join(wd, "crosspoints", "crosspoint_04.00"))
            txt = subprocess.check_output([os.path.join(O.REF_DIR, "dump_alignment"), os.path.join(wd, "alignment.00.bin")], text=True)
            dump = json.loads(txt[txt.index("{"):])          # the reference's reader prints "TRIM: ..." lines first
        out[name] = {
            "generator": {"m": m, "n": n, "segments": [list(s) for s in segs], "p_s": ps, "p_d": pd, "p_i": pi, "K": K, "seed": seed},
            "flags": extra,
            "seq_sha256": [hashlib.sha256(a.tobytes()).hexdigest(), hashlib.sha256(b.tobytes()).hexdigest()],
            "crosspoint_04": pts,
            "alignment": dump,
        }
        sizes = [max(q[1] - p[1], q[2] - p[2]) for p, q in zip(pts, pts[1:])]
        print(name, "crosspoints", len(pts), "largest partition", max(sizes), "score", dump["raw_score"],
              "gaps", len(dump["gaps0"]), len(dump["gaps1"]))
    with open(os.path.join(HERE, "stage5_runs.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"), sort_keys=True)


if __name__ == "__main__":
    main()
