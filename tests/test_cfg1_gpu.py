"""BASELINE config 1: 1 Mbp x 1 Mbp, SW stage-1 best score + end coordinate against the reference's CPU Gotoh path.

tests/golden/cfg1_stage1.json holds what the reference's own CPUBlockProcessor (oracle/_ref/oracle_cpu_block
--stage-1 --no-flush --fork=8, ~9 min on 8 cores; tests/golden/make_cfg1_golden.py) wrote for the cfg1 pair, one
crosspoint_01.00 per --fork column slice (1-based):
  * the LAST slice's entry is the best cell of the whole matrix (every process adds the best of the processes on its
    left to its own list, sw_stage1.cpp:420-426);
  * the entries of the other slices are the best cell ON THE LAST COLUMN of the slice (AlignerManager::dispatchColumn
    -> bestScoreLastColumn, AlignerManager.cpp:339-347; sw_stage1.cpp:230-236,480-486) -- seven independent probes of
    the matrix at columns 125000, 250000, ..., 875000, compared here with the last column of the prefix partition."""
import hashlib
import json
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import synth  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cfg1_stage1.json")))


@pytest.fixture(scope="module")
def cfg1():
    a, b = synth.make_config("cfg1")
    assert [hashlib.sha256(a.tobytes()).hexdigest(), hashlib.sha256(b.tobytes()).hexdigest()] == GOLD["seq_sha256"]
    return a, b


def _global_best():
    t, i, j, score = GOLD["slices"][-1]["crosspoint"]
    return (score, i - 1, j - 1)


@pytest.mark.parametrize("kernel,prune", [("s16x2", True), ("s16x2", False), ("s32", False)])
def test_cfg1_best_matches_reference_cpu(b200, cfg1, kernel, prune):
    a, b = cfg1
    al = b200.Aligner(kernel=b200.KERNEL_S32 if kernel == "s32" else b200.KERNEL_S16X2)
    al.set_sequences(a, b)
    r = al.align_partition(use_callbacks=False, prune=prune)
    assert r["best"] == _global_best() == (543605, 899999, 907656)
    if prune:
        assert r["cells"] < a.size * b.size
    al.close()


def test_cfg1_last_column_of_every_fork_slice(b200, cfg1):
    """Columns 125000 ... 875000 of the matrix: the reference's best-on-last-column cells of slices 0..6."""
    a, b = cfg1
    al = b200.Aligner(kernel=b200.KERNEL_S16X2)
    al.set_sequences(a, b)
    for s in GOLD["slices"][:-1]:
        t, i, j, score = s["crosspoint"]
        assert j == s["j1"]
        r = al.align_partition(0, 0, a.size, s["j1"], want_last_column=True, want_best_score=False)
        col = r["last_column"]                      # [0] = first-row cell, [i] = row i (1-based) of column j1
        assert col.size == a.size + 1
        assert int(col["h"].max()) == score, s["slice"]
        assert int(col["h"][i]) == score, s["slice"]
    al.close()


def test_cfg1_chain_matches_reference_cpu(b200, cfg1, monkeypatch):
    """Same pair through the block-cyclic chain (4 ranks sharing device 0), pruning on."""
    monkeypatch.setenv("B200_GROUP_WARPS_PER_SM", "4")
    monkeypatch.setenv("B200_WATCHDOG_S", "30")
    a, b = cfg1
    g = b200.Group([0, 0, 0, 0], a.size, b.size, 0)
    g.set_sequences(a, b)
    r = g.align_partition(use_callbacks=False, prune=True)
    assert r["best"] == _global_best()
    g.close()
