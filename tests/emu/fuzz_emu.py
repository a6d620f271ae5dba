"""TEST INFRASTRUCTURE: differential fuzzing of the library's device code on the SIMT emulation (tests/emu/) against the CPU
oracle.  Random partition shapes around the kernels' tile boundaries (32 / 64 / 512 / 1024), both kernels, SW / NW, every
border initialisation, caller-supplied borders in the middle of the sequences, N / IUPAC bytes, special rows, block pruning,
the chain with one handle and with groups of 2-4 ranks on distinct emulated devices, random chunk widths.

    python tests/emu/fuzz_emu.py [--cases N] [--seed S] [--max-side M] [--shuffle]

Prints one line per case and stops at the first mismatch with everything needed to replay it (--seed S --only K).
tests/test_emu_cpu.py runs a short, fixed-seed sweep; longer sweeps are run by hand (profiles/r02_emu_fuzz.txt)."""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402


def side(rng, max_side):
    """A length near a tile boundary, or log-uniform."""
    if rng.random() < 0.5:
        base = int(rng.choice([1, 2, 16, 31, 32, 33, 63, 64, 65, 127, 128, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049, 3072, 4096]))
        return max(1, min(max_side, base + int(rng.integers(-2, 3))))
    return int(min(max_side, max(1, round(2 ** rng.uniform(0, np.log2(max_side))))))


def make_case(rng, max_side):
    import synth
    m, n = side(rng, max_side), side(rng, max_side)
    if rng.random() < 0.15:                          # tall and narrow: the only shape with special rows (8192-row floor) that stays cheap
        m, n = int(rng.integers(8193, 20001)), int(min(n, rng.integers(1, 601)))
    lo, hi = sorted(rng.uniform(0, 1, 2))
    a, b = synth.make_pair(m, n, [(int(m * lo), int(m * hi))], 0.05, 0.02, 0.02, 0, int(rng.integers(1, 1 << 30)))
    c = dict(m=m, n=n, a=a, b=b)
    c["kernel"] = str(rng.choice(["s16x2", "s32"]))
    c["nw"] = bool(rng.random() < 0.4)
    c["mode"] = str(rng.choice(["plain", "plain", "self", "group"]))
    c["world"] = int(rng.integers(2, 5)) if c["mode"] == "group" else 1
    c["chunk"] = int(rng.choice([0, -1, 1, 31, 32, 33, 100, 257, 1000, 4096])) if c["mode"] != "plain" else 0
    if c["mode"] != "plain" and c["chunk"] > 0 and n // c["chunk"] > 400:
        c["chunk"] = max(c["chunk"], n // 400)      # keep the job count of a fuzz case small
    c["mixed"] = bool(rng.random() < 0.2)
    if c["mixed"]:
        for s in (a, b):
            k = int(rng.integers(0, max(1, s.size // 50) + 1))
            if k:
                s[rng.integers(0, s.size, k)] = rng.choice(np.frombuffer(b"NRYKMn", np.uint8), k)
    c["sub"] = bool(c["mode"] != "group" and rng.random() < 0.25 and m > 8 and n > 8)
    c["prune"] = bool(not c["nw"] and not c["sub"] and c["kernel"] == "s16x2" and not c["mixed"] and rng.random() < 0.35)
    c["special"] = bool(rng.random() < 0.5)
    if c["nw"]:
        c["rt"], c["ct"] = int(rng.integers(0, 3)), int(rng.integers(0, 3))     # INIT_ZEROES / INIT_GAPS / INIT_GAPS_OPENED
    return c


def run_case(b200, O, c):
    a, b, m, n = c["a"], c["b"], c["m"], c["n"]
    kern = b200.KERNEL_S32 if c["kernel"] == "s32" else b200.KERNEL_S16X2
    rec = b200.NEEDLEMAN_WUNSCH if c["nw"] else b200.SMITH_WATERMAN
    orec = O.NW if c["nw"] else O.SW
    kw = dict(recurrence=rec, want_last_row=True, want_last_column=True, want_best_score=True)
    okw = {}
    i0 = j0 = 0
    i1, j1 = m, n
    if c["sub"]:
        rng = np.random.default_rng(m * 7919 + n)
        i0, i1 = sorted(int(x) for x in rng.choice(np.arange(0, m + 1), 2, replace=False))
        j0, j1 = sorted(int(x) for x in rng.choice(np.arange(0, n + 1), 2, replace=False))
        fr = np.zeros(j1 - j0 + 1, O.CELL); fc = np.zeros(i1 - i0 + 1, O.CELL)
        fr["h"] = -np.cumsum(rng.integers(0, 4, fr.size)); fr["x"] = fr["h"] - rng.integers(1, 9, fr.size)
        fc["h"] = -np.cumsum(rng.integers(0, 4, fc.size)); fc["x"] = fc["h"] - rng.integers(1, 9, fc.size)
        if rng.random() < 0.4:                   # the lower right of a huge matrix: scores in the millions (s16 frame far from zero)
            off = int(rng.choice([100_000, 3_000_000, 400_000_000]))
            fr["h"] += off; fr["x"] += off; fc["h"] += off; fc["x"] += off
        fc[0] = fr[0]
        kw.update(first_row_init=b200.INIT_CUSTOM, first_col_init=b200.INIT_CUSTOM, first_row=fr, first_col=fc)
        okw.update(first_row=fr, first_col=fc)
    elif c["nw"]:
        kw.update(first_row_init=c["rt"], first_col_init=c["ct"])
        okw.update(first_row_type=c["rt"], first_col_type=c["ct"])
    if c["special"]:
        kw.update(want_special_rows=True, special_row_interval=1000)
    if c["prune"]:
        kw.update(prune=True)
    rows, cols = i1 - i0, j1 - j0

    if c["mode"] == "group":
        os.environ["B200_GROUP_WARPS_PER_SM"] = str(max(4, 16 // c["world"]))
        g = b200.Group(list(range(c["world"])), m, n, c["chunk"], kernel=kern)
        g.set_sequences(a, b)
        r = g.align_partition(mgpu=False, chunk_cols=c["chunk"], **kw)
        g.close()
    else:
        al = b200.Aligner(kernel=kern)
        if c["mode"] == "self":
            al.mgpu_setup(None, 0, 1, m, n, c["chunk"])
        al.set_sequences(a, b)
        r = al.align_partition(i0, j0, i1, j1, mgpu=c["mode"] == "self", chunk_cols=c["chunk"], **kw)
        al.close()

    ids = sorted(i for i in r["rows"] if i != i1)
    o = O.full_matrix(a[i0:i1], b[j0:j1], orec, row_ids=[i - i0 - 1 for i in ids] + [rows - 1], **okw)
    want_best = (o["best"][0], o["best"][1] + i0, o["best"][2] + j0) if o["best"][1] >= 0 else o["best"]
    if not c["nw"] or c["sub"]:
        assert r["best"] == want_best, ("best", r["best"], want_best)
    if c["prune"]:
        return "best only (pruned)"
    skip = 1 if c["sub"] else 0        # slot 0 of a custom border is the caller's corner cell
    for i in ids + [i1]:
        assert np.array_equal(r["rows"][i][skip:], o["rows"][i - i0 - 1][skip:]), ("row", i)
    assert np.array_equal(r["last_column"][skip:], o["last_col"][skip:]), "last column"
    assert r["cells"] == rows * cols, ("cells", r["cells"], rows * cols)
    return f"{len(ids)} special rows"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=50)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-side", type=int, default=5000)
    ap.add_argument("--only", type=int, default=-1, help="run just this case of the sweep")
    ap.add_argument("--shuffle", action="store_true", help="random fiber order inside every CTA (B200_EMU_SHUFFLE)")
    args = ap.parse_args()
    import build_emu
    os.environ["B200_LIB"] = build_emu.build()
    os.environ["B200_TEST_EMULATION"] = "1"
    os.environ.setdefault("B200_WATCHDOG_S", "30")
    if args.shuffle:
        os.environ["B200_EMU_SHUFFLE"] = str(args.seed + 100)
    from __graft_entry__ import load_package
    import oracle_lib as O
    b200 = load_package()
    assert b200.load_library().b200_emu_is_emulation() == 1
    bad = 0
    import ctypes
    ovf = b200.load_library().b200_emu_s16_overflows
    ovf.restype = ctypes.c_longlong
    for k in range(args.cases):
        rng = np.random.default_rng([args.seed, k])
        c = make_case(rng, args.max_side)
        if args.only >= 0 and k != args.only:
            continue
        tag = (f"case {k}: {c['m']}x{c['n']} {c['kernel']} {'NW' if c['nw'] else 'SW'} mode={c['mode']} world={c['world']} chunk={c['chunk']} "
               f"sub={int(c['sub'])} mixed={int(c['mixed'])} prune={int(c['prune'])} special={int(c['special'])}")
        try:
            o0 = ovf()
            what = run_case(b200, O, c)
            print(tag, "ok:", what, f"[{ovf() - o0} s16 overflow events]" if ovf() != o0 else "", flush=True)
        except Exception as e:                      # noqa: BLE001
            bad += 1
            print(tag, "MISMATCH:", repr(e)[:400], f"   replay: --seed {args.seed} --only {k}", flush=True)
            break
    print(f"fuzz: {args.cases if args.only < 0 else 1} cases, {bad} mismatches, {ovf()} s16 overflow events (seed {args.seed})")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
