"""BASELINE config 1: 1 Mbp x 1 Mbp, SW stage-1 best score + end coordinate against the reference's CPU Gotoh path.

tests/golden/cfg1_stage1.json holds what the reference's own CPUBlockProcessor (oracle/_ref/oracle_cpu_block
--stage-1 --no-flush --fork=8, ~9 min on 8 cores; tests/golden/make_cfg1_golden.py) wrote for the cfg1 pair: the best
cell of each of the eight --fork column slices (sw_stage1.cpp:480-491, 1-based).  The best over columns [0, j1_k) is the
maximum of the first k+1 entries (canonical tie-break), which each kernel must reproduce with and without pruning."""
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


def _prefix_best(k):
    cand = [(s["crosspoint"][3], s["crosspoint"][1] - 1, s["crosspoint"][2] - 1) for s in GOLD["slices"][:k + 1]]
    return max(cand, key=lambda c: (c[0], -c[1], -c[2]))


@pytest.mark.parametrize("kernel,prune", [("s16x2", True), ("s16x2", False), ("s32", False)])
def test_cfg1_best_matches_reference_cpu(b200, cfg1, kernel, prune):
    a, b = cfg1
    al = b200.Aligner(kernel=b200.KERNEL_S32 if kernel == "s32" else b200.KERNEL_S16X2)
    al.set_sequences(a, b)
    r = al.align_partition(use_callbacks=False, prune=prune)
    assert r["best"] == _prefix_best(7) == (543605, 899999, 907656)
    if prune:
        assert r["cells"] < a.size * b.size
    al.close()


def test_cfg1_every_fork_slice(b200, cfg1):
    """Prefix partitions [0, j1_k): the reference's per-slice bests, one by one (packed kernel, pruning on)."""
    a, b = cfg1
    al = b200.Aligner(kernel=b200.KERNEL_S16X2)
    al.set_sequences(a, b)
    for k, s in enumerate(GOLD["slices"]):
        r = al.align_partition(0, 0, a.size, s["j1"], use_callbacks=False, prune=True, super_i1=a.size, super_j1=s["j1"])
        assert r["best"] == _prefix_best(k), k
    al.close()


def test_cfg1_chain_matches_reference_cpu(b200, cfg1, monkeypatch):
    """Same pair through the block-cyclic chain (4 ranks sharing device 0), pruning on."""
    monkeypatch.setenv("B200_GROUP_WARPS_PER_SM", "4")
    monkeypatch.setenv("B200_WATCHDOG_S", "30")
    a, b = cfg1
    g = b200.Group([0, 0, 0, 0], a.size, b.size, 0)
    g.set_sequences(a, b)
    r = g.align_partition(use_callbacks=False, prune=True)
    assert r["best"] == (543605, 899999, 907656)
    g.close()
