"""CPU tests (no GPU): pin the plain-C oracle (oracle/gotoh_oracle.c) against the reference.

 * tests/golden/reference_runs.json was produced by the reference's own code (oracle/_ref/oracle_cpu, see
   tests/golden/make_golden.py): stage-1 best cell and sha256 of every stage-1 special-row file.
 * when the reference-built binaries are present (build container), the Block-family binary is executed too."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import synth  # noqa: E402

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_runs.json")))


def _pair(g):
    a, b = synth.make_pair(g["m"], g["n"], [tuple(g["homology"])], g["p_s"], g["p_d"], g["p_i"], 0, g["seed"])
    return a, b


@pytest.mark.parametrize("name", sorted(GOLD))
def test_generator_is_reproducible(name):
    a, b = _pair(GOLD[name]["generator"])
    assert [hashlib.sha256(a.tobytes()).hexdigest(), hashlib.sha256(b.tobytes()).hexdigest()] == GOLD[name]["seq_sha256"]


@pytest.mark.parametrize("name", ["sw_3k", "sw_12k_rows", "sw_40k"])
def test_sw_best_and_special_rows_match_reference(name):
    g = GOLD[name]
    a, b = _pair(g["generator"])
    ids = sorted(int(os.path.basename(k), 16) for k in g["special_rows_stage1"])
    o = O.full_matrix(a, b, O.SW, row_ids=[i - 1 for i in ids], want_last_col=False)
    (_t, i, j, score), = g["crosspoints"]["crosspoint_01.00"]
    assert o["best"] == (score, i - 1, j - 1)            # crosspoint files are 1-based (AlignerManager.cpp:411-415)
    for k, meta in g["special_rows_stage1"].items():
        rid = int(os.path.basename(k), 16)
        row = o["rows"][rid - 1]
        assert row.size == meta["cells"]
        assert [int(row[0]["h"]), int(row[0]["x"])] == meta["first_cell"]
        assert hashlib.sha256(row.tobytes()).hexdigest() == meta["sha256"], f"{name}: special row {rid:#x}"


def test_nw_global_matches_reference():
    g = GOLD["nw_20k"]
    a, b = _pair(g["generator"])
    ids = sorted(int(os.path.basename(k), 16) for k in g["special_rows_stage1"])
    m, n = a.size, b.size
    o = O.full_matrix(a, b, O.NW, first_row_type=O.INIT_GAPS, first_col_type=O.INIT_GAPS, row_ids=[i - 1 for i in ids] + [m - 1])
    (_t, i, j, score), = g["crosspoints"]["crosspoint_01.00"]
    assert (i, j) == (m, n) and int(o["rows"][m - 1][n]["h"]) == score
    for k, meta in g["special_rows_stage1"].items():
        rid = int(os.path.basename(k), 16)
        assert hashlib.sha256(o["rows"][rid - 1].tobytes()).hexdigest() == meta["sha256"], f"special row {rid:#x}"


def test_block_family_binary_agrees(tmp_path):
    if not O.have_ref_binaries():
        pytest.skip("oracle/_ref binaries not built")
    g = GOLD["sw_3k"]
    a, b = _pair(g["generator"])
    fa, fb = str(tmp_path / "A.fa"), str(tmp_path / "B.fa")
    synth.write_fasta(fa, a, "A"); synth.write_fasta(fb, b, "B")
    wd = O.run_ref("oracle_cpu_block", fa, fb, str(tmp_path / "w"), ["--stage-1", "--no-flush"])
    pts = O.read_crosspoints(os.path.join(wd, "crosspoints", "crosspoint_01.00"))
    o = O.full_matrix(a, b, O.SW, want_last_col=False)
    assert pts == [(0, o["best"][1] + 1, o["best"][2] + 1, o["best"][0])]


def test_match_column_rules():
    # first k wins; match (H+H) before gap (E+E+open) at the same k; overshoot is an error (AlignerUtils.cpp:59-84)
    buf = np.zeros(6, O.CELL); base = np.zeros(6, O.CELL)
    buf["h"] = [1, 2, 3, 4, 5, 6]; base["h"] = [0, 0, 0, 6, 0, 4]
    buf["x"] = -100; base["x"] = -100
    r = O.match_column(buf, base, 10)
    assert r == dict(found=True, k=3, score=6, type=0)
    buf["x"][1] = 3; base["x"][1] = 4                     # 3 + 4 + 3 == 10 at k=1 -> gapped match wins (earlier k)
    r = O.match_column(buf, base, 10)
    assert r == dict(found=True, k=1, score=4, type=1)
    base["h"][0] = 20                                      # overshoot at k=0
    r = O.match_column(buf, base, 10)
    assert not r["found"] and r["type"] == -1


def test_init_cells():
    z = O.init_cells(4, O.INIT_ZEROES)
    assert list(z["h"]) == [0, 0, 0, 0] and set(z["x"]) == {-O.INF}
    g = O.init_cells(4, O.INIT_GAPS)
    assert list(g["h"]) == [0, -5, -7, -9]
    g = O.init_cells(4, O.INIT_GAPS_OPENED)
    assert list(g["h"]) == [0, -2, -4, -6]


@pytest.mark.parametrize("name", sorted(GOLD))
def test_stage4_restatement_matches_reference(name):
    """ort_split_2 + split_thread + merge_partitions restated in C reproduce the reference's crosspoint_04 file."""
    g = GOLD[name]
    a, b = _pair(g["generator"])
    src = sorted(k for k in g["crosspoints"] if k.startswith("crosspoint_03"))[-1]
    pts = O.stage4(a, b, O.golden_points(g["crosspoints"][src]), 16)
    assert np.array_equal(pts, O.golden_points(g["crosspoints"]["crosspoint_04.00"]))


S5_GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "stage5_runs.json")))


def s5_pair(name):
    ge = S5_GOLD[name]["generator"]
    a, b = synth.make_pair(ge["m"], ge["n"], [tuple(s) for s in ge["segments"]], ge["p_s"], ge["p_d"], ge["p_i"], ge["K"], ge["seed"])
    assert [hashlib.sha256(a.tobytes()).hexdigest(), hashlib.sha256(b.tobytes()).hexdigest()] == S5_GOLD[name]["seq_sha256"]
    return a, b


@pytest.mark.parametrize("name", sorted(S5_GOLD))
def test_stage5_restatement_matches_reference(name):
    """go_stage5 walked over the reference's crosspoint_04 reproduces what the reference's stage 5 stored in
    alignment.00.bin (read back with the reference's own reader, tests/golden/make_stage5_golden.py): raw score, the
    four counters of total_score_t and both gap lists."""
    g = S5_GOLD[name]
    a, b = s5_pair(name)
    pts = O.golden_points(g["crosspoint_04"])
    ops, off, ln, st = O.stage5(a, b, pts)
    al = g["alignment"]
    assert st["score"] == al["raw_score"] == int(pts["score"][-1] - pts["score"][0])
    assert [st[k] for k in ("matches", "mismatches", "gap_open", "gap_ext")] == [al[k] for k in ("matches", "mismatches", "gap_open", "gap_ext")]
    g0, g1 = O.stage5_gaps(pts, ops, off, ln)
    assert g0 == sorted(al["gaps0"]) and g1 == sorted(al["gaps1"])
    assert al["start"] == [int(pts["i"][0]) + 1, int(pts["j"][0]) + 1] and al["end"] == [int(pts["i"][-1]), int(pts["j"][-1])]


@pytest.mark.parametrize("seed", range(8))
def test_stage5_walk_scores_the_global_optimum(seed):
    """Two independent restatements must agree: the score the stage-5 walk collects over a MATCH -> MATCH partition
    (sw_stage5.cpp tables + traceback) is the Needleman-Wunsch optimum of the two substrings, i.e. the last cell of the
    stage-1 recurrence (CPUBlockProcessor cell + InitialCellsReader gap borders).  Also: the walk consumes the partition
    exactly, and its counters add up to its length."""
    rng = np.random.default_rng(300 + seed)
    m, n = int(rng.integers(1, 400)), int(rng.integers(1, 400))
    a = synth.ACGT[rng.integers(0, 4, size=m)]
    b = a.copy()[:n] if seed % 2 and n <= m else synth.ACGT[rng.integers(0, 4, size=n)]
    if seed % 2:                                       # related sequences: sprinkle substitutions
        b = b.copy()
        hit = rng.random(b.size) < 0.2
        b[hit] = synth.ACGT[rng.integers(0, 4, size=int(hit.sum()))]
    n = b.size
    pts = np.array([(0, 0, 0, 0), (m, n, 0, 0)], dtype=O.XPOINT)
    ops, off, ln, st = O.stage5(a, b, pts)
    nw = O.full_matrix(a, b, O.NW, first_row_type=O.INIT_GAPS, first_col_type=O.INIT_GAPS)
    assert st["score"] == int(nw["last_col"][-1]["h"])
    walk = ops[:int(ln[1])]
    assert int((walk == 0).sum() + (walk == 1).sum()) == m and int((walk == 0).sum() + (walk == 2).sum()) == n
    assert st["matches"] + st["mismatches"] == int((walk == 0).sum()) and st["gap_ext"] == int((walk != 0).sum())
    # the reference charges the opening of a gap that runs into the partition's border to the score without counting it in
    # gapOpen (sw_stage5.cpp:291-311), so the counters explain the score up to that one opening
    assert st["score"] - (st["matches"] - 3 * st["mismatches"] - 3 * st["gap_open"] - 2 * st["gap_ext"]) in (0, -3)
