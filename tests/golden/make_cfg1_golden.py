#!/usr/bin/env python3
"""Regenerates tests/golden/cfg1_stage1.json from the REFERENCE's own CPU path (BASELINE config 1).

BASELINE.json config 1 = "1 Mbp x 1 Mbp synthetic DNA pair, SW stage-1 best score + end coordinate vs the CPU Gotoh
path".  The CPU Gotoh path is the reference's CPUBlockProcessor (C/libmasa/processors/CPUBlockProcessor.cpp:66-174)
driven by MASA-Core's own --fork (C/libmasa/libmasa.cpp:540-642), i.e. oracle/_ref/oracle_cpu_block:

    oracle_cpu_block --stage-1 --no-flush --fork=8 cfg1_A.fa cfg1_B.fa        (about 9 minutes on 8 cores)

--fork splits seq1 into 8 equal column slices (libmasa.cpp:632-635); every forked process writes one crosspoint to
FORK.0k/crosspoints/crosspoint_01.00 (sw_stage1.cpp:480-491), 1-based: the LAST process the best cell of the whole
matrix (the bests travel rightwards, sw_stage1.cpp:420-426), every other process the best cell on the LAST COLUMN of
its slice (bestScoreLastColumn, AlignerManager.cpp:339-347; sw_stage1.cpp:230-236).  The fixture keeps all eight: the
global best plus seven probes of single matrix columns.  Needs oracle/_ref (build container only)."""
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

FORK = 8


def main():
    a, b = synth.make_config("cfg1")
    with tempfile.TemporaryDirectory() as td:
        fa, fb = os.path.join(td, "cfg1_A.fa"), os.path.join(td, "cfg1_B.fa")
        synth.write_fasta(fa, a, "synth_cfg1_A"); synth.write_fasta(fb, b, "synth_cfg1_B")
        wd = os.path.join(td, "w")
        exe = os.path.join(ROOT, "oracle", "_ref", "oracle_cpu_block")
        subprocess.run([exe, f"--work-dir={wd}", "--clear", "--verbose=0", "--stage-1", "--no-flush", f"--fork={FORK}", fa, fb],
                       check=True, cwd=td, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        slices = []
        for k in range(FORK):
            (t, i, j, s), = O.read_crosspoints(os.path.join(wd, f"FORK.{k:02d}", "crosspoints", "crosspoint_01.00"))
            j0, j1 = b.size * k // FORK, b.size * (k + 1) // FORK
            slices.append({"slice": k, "j0": j0, "j1": j1, "crosspoint": [t, i, j, s]})
    out = {
        "config": "cfg1", "generator": dict(synth.CONFIGS["cfg1"]),
        "seq_sha256": [hashlib.sha256(a.tobytes()).hexdigest(), hashlib.sha256(b.tobytes()).hexdigest()],
        "command": f"oracle/_ref/oracle_cpu_block --stage-1 --no-flush --fork={FORK}",
        "slices": slices,
    }
    with open(os.path.join(HERE, "cfg1_stage1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(slices))


if __name__ == "__main__":
    main()
