"""GPU parity of the batched stage-5 traceback (b200_stage5, csrc/stage5.cuh) through the C ABI:
  * the reference's own stage 5: golden crosspoint_04 -> the gap lists / counters stored in the reference's
    alignment.00.bin (tests/golden/stage5_runs.json, made by tests/golden/make_stage5_golden.py),
  * the oracle restatement (oracle/gotoh_oracle.c go_stage5, pinned to the same vectors) step for step on random
    partition chains (all type pairs, pure-gap partitions, both storage variants) and on a chain of 125,000 partitions.
The end-to-end proof -- alignment.00.bin / alignment.00.txt of build/cudalign byte-identical to the reference's -- is
tests/test_pipeline_gpu.py, whose binary runs this stage 5."""
import json
import os
import sys
import time

import numpy as np
import pytest

import oracle_lib as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import synth  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "stage5_runs.json")))


def same_walk(b200, a, b, pts):
    al = b200.Aligner()
    al.set_sequences(a, b)
    t0 = time.perf_counter()
    ops, off, ln, st = al.stage5(pts)
    dt = time.perf_counter() - t0
    launches = al.kernel_launches()
    al.close()
    ops_o, off_o, len_o, st_o = O.stage5(a, b, pts)
    assert np.array_equal(off, off_o)
    assert np.array_equal(ln, len_o)
    assert st == st_o
    lens = ln.astype(np.int64)
    pos = np.repeat(off - (np.cumsum(lens) - lens), lens) + np.arange(int(lens.sum()))       # every used slot
    if not np.array_equal(ops[pos], ops_o[pos]):
        bad = int(pos[np.nonzero(ops[pos] != ops_o[pos])[0][0]])
        k = int(np.searchsorted(off, bad, side="right")) - 1
        s = slice(int(off[k]), int(off[k]) + int(ln[k]))
        raise AssertionError(f"partition {k} {pts[k - 1]} -> {pts[k]}: {ops[s][:40]} != {ops_o[s][:40]}")
    assert launches >= 1
    return ops, off, ln, st, dt


@pytest.mark.parametrize("name", sorted(GOLD))
def test_stage5_reproduces_reference_alignment(b200, name):
    g = GOLD[name]
    ge = g["generator"]
    a, b = synth.make_pair(ge["m"], ge["n"], [tuple(s) for s in ge["segments"]], ge["p_s"], ge["p_d"], ge["p_i"], ge["K"], ge["seed"])
    pts = O.golden_points(g["crosspoint_04"])
    ops, off, ln, st, _dt = same_walk(b200, a, b, pts)
    al = g["alignment"]
    assert st["score"] == al["raw_score"]
    assert [st[k] for k in ("matches", "mismatches", "gap_open", "gap_ext")] == [al[k] for k in ("matches", "mismatches", "gap_open", "gap_ext")]
    g0, g1 = O.stage5_gaps(pts, ops, off, ln)
    assert g0 == sorted(al["gaps0"]) and g1 == sorted(al["gaps1"])


@pytest.mark.parametrize("seed", range(4))
def test_stage5_random_partition_chains(b200, seed):
    """Chains that are NOT optimal paths: every start/end type pair, sides 0..90 (so the thread-local and the global-scratch
    variants both run, and pure-gap partitions are interleaved), tie-rich two-letter sequences on odd seeds."""
    rng = np.random.default_rng(2000 + seed)
    m = n = 20000
    a, b = synth.make_pair(m, n, [(0, m)], 0.10, 0.05, 0.05, 0, 700 + seed)
    if seed % 2:
        a = synth.ACGT[rng.integers(0, 2, size=m)]
        b = synth.ACGT[rng.integers(0, 2, size=n)]
    pts = [(0, 0, int(rng.integers(0, 3)), 0)]
    while True:
        i, j = pts[-1][0], pts[-1][1]
        di, dj = int(rng.integers(0, 91)), int(rng.integers(0, 91))
        if rng.random() < 0.6:
            di, dj = min(di, 16), min(dj, 16)
        if i + di > m or j + dj > n:
            break
        if di == 0 and dj == 0:
            continue
        pts.append((i + di, j + dj, int(rng.integers(0, 3)), 0))
    same_walk(b200, a, b, np.array(pts, dtype=O.XPOINT))


def test_stage5_large_partitions(b200):
    """--maximum-partition up to 1024: a few partitions of several hundred cells on a side (global-scratch variant)."""
    rng = np.random.default_rng(5)
    m, n = 9000, 9500
    a, b = synth.make_pair(m, n, [(0, m)], 0.05, 0.02, 0.02, 0, 77)
    pts = [(0, 0, 0, 0)]
    for di, dj in [(1000, 1024), (513, 700), (1, 900), (1024, 1), (333, 333), (900, 1000), (1024, 1024)]:
        pts.append((pts[-1][0] + di, pts[-1][1] + dj, int(rng.integers(0, 3)), 0))
    same_walk(b200, a, b, np.array(pts, dtype=O.XPOINT))


def test_stage5_many_partitions(b200):
    """125,000 partitions of 16 x 16 (the default --maximum-partition) along the diagonal of a 2M x 2M pair: what stage 5 of
    a chromosome-sized alignment looks like.  Same walk as the oracle; the time is printed for the record."""
    m = n = 2_000_000
    a, b = synth.make_pair(m, n, [(0, m)], 0.03, 0.0, 0.0, 0, 31)          # substitutions only: the diagonal is the path
    k = np.arange(0, m // 16 + 1, dtype=np.int64)
    pts = np.zeros(k.size, dtype=O.XPOINT)
    pts["i"] = pts["j"] = 16 * k
    _ops, _off, _ln, st, dt = same_walk(b200, a, b, pts)
    print(f"stage 5: {k.size - 1} partitions of 16x16 in {dt * 1e3:.1f} ms (includes planning, H2D, D2H of {2 * m} step bytes); "
          f"matches {st['matches']} mismatches {st['mismatches']}")
    assert st["matches"] > 0.9 * m
