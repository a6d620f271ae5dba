"""TEST INFRASTRUCTURE: differential fuzzing of the drop-in binary on the SIMT emulation.  Random FASTA pairs and MASA-Core
flags through build/cudalign (B200Aligner + the reference's unmodified MASA-Core + the stage-4 / stage-5 substitutes, with
the emulation library pre-loaded in place of libb200align.so) and through oracle/_ref/oracle_cpu (the reference's CPU run);
every artefact must be identical: crosspoint files of stages 1-4, alignment.00.bin, alignment.00.txt, special-row files.

    python tests/emu/fuzz_pipeline.py [--cases N] [--seed S] [--max-side M]

Stops at the first difference and prints how to replay it (--seed S --only K)."""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=20)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-side", type=int, default=20000)
    ap.add_argument("--only", type=int, default=-1)
    args = ap.parse_args()
    import build_emu
    import synth
    import test_pipeline_gpu as P
    emu = build_emu.build()
    cudalign = os.path.join(ROOT, "build", "cudalign")
    oracle = os.path.join(ROOT, "oracle", "_ref", "oracle_cpu")
    assert os.path.exists(cudalign) and os.path.exists(oracle), "build/cudalign and oracle/_ref/oracle_cpu are needed (python __graft_entry__.py)"
    bad = 0
    for k in range(args.cases):
        rng = np.random.default_rng([args.seed, k, 77])
        m = int(round(2 ** rng.uniform(np.log2(300), np.log2(args.max_side))))
        n = int(max(50, m * rng.uniform(0.5, 1.5)))
        lo, hi = sorted(rng.uniform(0, 1, 2))
        if hi - lo < 0.2:
            lo, hi = 0.1, 0.9
        a, b = synth.make_pair(m, n, [(int(m * lo), int(m * hi))], float(rng.choice([0.02, 0.05, 0.1])), 0.02, 0.02, int(rng.integers(0, 3)), int(rng.integers(1, 1 << 30)))
        if rng.random() < 0.25:                  # N / IUPAC / lower-case bytes: the byte-compare semantics of the reference (mixed kernels)
            for sq in (a, b):
                cnt = int(rng.integers(1, max(2, sq.size // 40)))
                sq[rng.integers(0, sq.size, cnt)] = rng.choice(np.frombuffer(b"NNNRYKMacgtn", np.uint8), cnt)
                if sq.size > 400:
                    p0 = int(rng.integers(0, sq.size - 300))
                    sq[p0:p0 + int(rng.integers(10, 300))] = ord("N")
        extra = []
        if rng.random() < 0.15:
            extra.append("--clear-n")
        r_edges = rng.random()
        nw = r_edges < 0.3
        if nw:
            extra += ["--alignment-edges=++"]
        elif r_edges < 0.5:                      # one of the other 23 start/end rules (semi-global kinds, libmasa.cpp:258-264)
            ed = "**"
            while ed in ("**", "++"):
                ed = str(rng.choice(list("*123+"))) + str(rng.choice(list("*123+")))
            extra += [f"--alignment-edges={ed}"]
            nw = True                            # (pruning is a stage-1 SW-local feature: keep it off for these)
        if nw or rng.random() < 0.5:
            extra += ["--no-block-pruning"]
        sra = str(rng.choice(["", "--ram-size=1M", "--disk-size=2M", "--ram-size=200K"]))
        if sra:
            extra.append(sra)
        if rng.random() < 0.3:
            extra.append(f"--maximum-partition={int(rng.choice([32, 64, 200]))}")
        new_extra, env = [], {"B200_WATCHDOG_S": "30"}
        mode = str(rng.choice(["fast", "fast", "diag", "gpus2", "gpus3"]))
        if mode == "diag":
            new_extra = ["--no-fast-path"]
        elif mode.startswith("gpus"):
            w = int(mode[4:])
            new_extra = ["--gpus=" + ",".join(str(d) for d in range(w))]
            env.update({"B200_GROUP_WARPS_PER_SM": str(16 // 4 if w == 3 else 16 // w), "B200_GROUP_MIN_CELLS": "0", "B200_CHAIN_CHUNK": str(int(rng.choice([300, 1000, 3000])))})
        if args.only >= 0 and k != args.only:
            continue
        tag = f"case {k}: {m}x{n} {' '.join(extra)} [{mode}]"
        td = tempfile.mkdtemp(prefix="b200fuzz_")
        try:
            fa, fb = os.path.join(td, "A.fa"), os.path.join(td, "B.fa")
            synth.write_fasta(fa, a, "A"); synth.write_fasta(fb, b, "B")
            w_ref, w_new = os.path.join(td, "ref"), os.path.join(td, "new")
            P._run(oracle, fa, fb, w_ref, extra)
            env["LD_PRELOAD"] = emu
            P._run(cudalign, fa, fb, w_new, extra + new_extra, env=env)
            sr = "--no-block-pruning" in extra and bool(sra)       # special-row files are comparable with pruning off
            if not os.path.exists(os.path.join(w_ref, "alignment.00.bin")):
                # the reference itself ends without an alignment (an empty optimum of a semi-global mode): so must the binary
                assert not os.path.exists(os.path.join(w_new, "alignment.00.bin")), "the reference wrote no alignment, build/cudalign did"
                xr, xn = P._files(os.path.join(w_ref, "crosspoints")), P._files(os.path.join(w_new, "crosspoints"))
                assert sorted(xr) == sorted(xn) and all(open(xr[q]).read() == open(xn[q]).read() for q in xr), "crosspoint files differ"
                print(tag, "ok (no alignment on either side; crosspoint files identical)", flush=True)
                continue
            nrows = P._compare(w_ref, w_new, sr)
            print(tag, f"ok ({nrows} special-row files compared)", flush=True)
        except (AssertionError, OSError, subprocess.SubprocessError) as e:
            bad += 1
            keep = tempfile.mkdtemp(prefix="b200fuzz_failed_")
            with open(os.path.join(keep, "output.txt"), "w") as f:
                f.write(str(e))
            for nm in ("A.fa", "B.fa"):
                shutil.copy(os.path.join(td, nm), keep)
            print(tag, "DIFFERENT:", str(e)[-2500:], f"\n   inputs and full output kept in {keep}; replay: --seed {args.seed} --only {k}", flush=True)
            break
        finally:
            shutil.rmtree(td, ignore_errors=True)
    print(f"pipeline fuzz: {bad} differences (seed {args.seed})")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
