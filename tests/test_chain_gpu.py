"""GPU parity of the block-cyclic chain (b200_align_partition with B200_MGPU_CHAIN, b200_group_*) on ONE GPU.

The chain is the multi-GPU path: column chunks dealt round-robin to the GPUs, borders handed over
through the exchange blocks, jobs scheduled by left/top events.  Nothing in that machinery needs the ranks to sit on
different devices, so the single-GPU test box exercises all of it:
  * world = 1: the GPU is its own neighbour (every chunk border goes through the exchange block);
  * world = 2 / 4: several handles of one b200_group on device 0, each limited to a share of the SMs' warp slots so
    that their persistent kernels are co-resident (the same code runs with one handle per GPU on an NVLink box:
    tests/test_mgpu_gpu.py).
Everything is compared bit-exactly with the CPU oracle: best cell, special rows, last row, last column, with and
without block pruning, SW and NW, both kernels, ragged chunk widths, repeated calls on the same handles."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import synth  # noqa: E402

pytestmark = pytest.mark.gpu


def _pair(m, n, seed, lo=0.1, hi=0.9, ps=0.05, pd=0.01, pi=0.01):
    return synth.make_pair(m, n, [(int(m * lo), int(m * hi))], ps, pd, pi, 0, seed)


def _kern(b200, name):
    return b200.KERNEL_S32 if name == "s32" else b200.KERNEL_S16X2


@pytest.fixture()
def short_watchdog(monkeypatch):
    monkeypatch.setenv("B200_WATCHDOG_S", "20")      # a protocol bug must fail the test, not sit out the default limit


@pytest.mark.parametrize("kernel", ["s16x2", "s32"])
@pytest.mark.parametrize("m,n,chunk", [(5000, 7000, 1000), (2049, 1500, 257), (20000, 9000, 4096), (700, 900, -1), (3000, 3000, 1), (1, 1, 0)])
def test_self_chain_sw_matches_oracle(b200, short_watchdog, kernel, m, n, chunk):
    a, b = _pair(m, n, 7)
    al = b200.Aligner(kernel=_kern(b200, kernel))
    al.mgpu_setup(None, 0, 1, m, n, chunk)
    al.set_sequences(a, b)
    for rep in range(2):                             # twice: the exchange block must re-arm
        r = al.align_partition(want_last_row=True, want_last_column=True, want_special_rows=True, special_row_interval=1000,
                               mgpu=True, chunk_cols=chunk)
        ids = sorted(i for i in r["rows"] if i != m)
        assert ids == list(range(8192, m, 8192))
        o = O.full_matrix(a, b, O.SW, row_ids=[i - 1 for i in ids] + [m - 1])
        assert r["best"] == o["best"], rep
        assert r["cells"] == m * n
        for i in ids + [m]:
            assert np.array_equal(r["rows"][i], o["rows"][i - 1]), f"row {i} (rep {rep})"
        assert np.array_equal(r["last_column"], o["last_col"])
    al.close()


@pytest.mark.parametrize("m,n,chunk", [(60000, 60000, 3000), (50000, 70000, 1500), (70000, 40000, 8192)])
def test_self_chain_pruning_keeps_best_exact(b200, short_watchdog, m, n, chunk):
    """Block pruning across chunk borders: the alignment path enters most chunks through the LOWER rows of a strip
    (lanes that start late), the case the left-border term of the pruning test exists for."""
    a, b = _pair(m, n, 23, 0.05, 0.95)
    o = O.full_matrix(a, b, O.SW, want_last_col=False)
    al = b200.Aligner(kernel=b200.KERNEL_S16X2)
    al.mgpu_setup(None, 0, 1, m, n, chunk)
    al.set_sequences(a, b)
    r = al.align_partition(mgpu=True, chunk_cols=chunk, prune=True, use_callbacks=False)
    assert r["best"] == o["best"]
    assert r["cells"] < 0.8 * m * n, "pruning did not engage: the test would prove nothing"
    r2 = al.align_partition(mgpu=True, chunk_cols=chunk, prune=False, use_callbacks=False)
    assert r2["best"] == o["best"] and r2["cells"] == m * n
    al.close()


def test_self_chain_rearm_with_other_sequences(b200, short_watchdog):
    """Second chained call on the same handle with a pair whose best is far LOWER than the first one's: a running best
    left over from the first call would over-prune (re-arm race on the shared best word)."""
    m, n = 40000, 40000
    al = b200.Aligner(kernel=b200.KERNEL_S16X2)
    al.mgpu_setup(None, 0, 1, m, n, 2000)
    for seed, lo, hi in ((5, 0.05, 0.95), (6, 0.45, 0.55), (7, 0.3, 0.6)):
        a, b = _pair(m, n, seed, lo, hi)
        al.set_sequences(a, b)
        r = al.align_partition(mgpu=True, chunk_cols=2000, prune=True, use_callbacks=False)
        assert r["best"] == O.full_matrix(a, b, O.SW, want_last_col=False)["best"], seed
    al.close()


@pytest.mark.parametrize("kernel", ["s16x2", "s32"])
@pytest.mark.parametrize("rt,ct", [(O.INIT_GAPS, O.INIT_GAPS), (O.INIT_GAPS_OPENED, O.INIT_GAPS), (O.INIT_GAPS, O.INIT_ZEROES)])
def test_self_chain_nw_global(b200, short_watchdog, kernel, rt, ct):
    m, n = 6001, 5500
    a, b = _pair(m, n, 21, 0.0, 1.0)
    al = b200.Aligner(kernel=_kern(b200, kernel))
    al.mgpu_setup(None, 0, 1, m, n, 700)
    al.set_sequences(a, b)
    r = al.align_partition(recurrence=b200.NEEDLEMAN_WUNSCH, first_row_init=rt, first_col_init=ct, want_last_row=True,
                           want_last_column=True, want_best_score=False, mgpu=True, chunk_cols=700)
    o = O.full_matrix(a, b, O.NW, first_row_type=rt, first_col_type=ct, row_ids=[m - 1])
    assert np.array_equal(r["rows"][m], o["rows"][m - 1])
    assert np.array_equal(r["last_column"], o["last_col"])
    al.close()


def test_self_chain_custom_borders_subpartition(b200, short_watchdog):
    a, b = _pair(9000, 9000, 31)
    i0, j0, i1, j1 = 700, 300, 8100, 7333
    rng = np.random.default_rng(5)
    fr = np.zeros(j1 - j0 + 1, O.CELL); fc = np.zeros(i1 - i0 + 1, O.CELL)
    fr["h"] = -np.cumsum(rng.integers(0, 4, fr.size)); fr["x"] = fr["h"] - rng.integers(1, 9, fr.size)
    fc["h"] = -np.cumsum(rng.integers(0, 4, fc.size)); fc["x"] = fc["h"] - rng.integers(1, 9, fc.size)
    fc[0] = fr[0]
    o = O.full_matrix(a[i0:i1], b[j0:j1], O.NW, first_row=fr, first_row_type=O.INIT_CUSTOM, first_col=fc,
                      first_col_type=O.INIT_CUSTOM, row_ids=[i1 - i0 - 1])
    for kernel in (b200.KERNEL_S32, b200.KERNEL_S16X2):
        al = b200.Aligner(kernel=kernel)
        al.mgpu_setup(None, 0, 1, i1 - i0, j1 - j0, 900)
        al.set_sequences(a, b)
        r = al.align_partition(i0, j0, i1, j1, recurrence=b200.NEEDLEMAN_WUNSCH, first_row_init=b200.INIT_CUSTOM,
                               first_col_init=b200.INIT_CUSTOM, first_row=fr, first_col=fc, want_last_row=True,
                               want_last_column=True, want_best_score=False, mgpu=True, chunk_cols=900)
        assert np.array_equal(r["rows"][i1], o["rows"][i1 - i0 - 1])
        assert np.array_equal(r["last_column"], o["last_col"])
        al.close()


# ---- several ranks on one device (b200_group with a share of the warp slots each) ------------------------------------
def _group(b200, world, m, n, chunk, kernel):
    g = b200.Group([0] * world, m, n, chunk, kernel=kernel)
    return g


@pytest.mark.parametrize("world,chunk", [(2, 1500), (4, 700), (2, -1), (3, 2048)])
@pytest.mark.parametrize("kernel", ["s16x2", "s32"])
def test_group_on_one_device_sw(b200, short_watchdog, monkeypatch, world, chunk, kernel):
    monkeypatch.setenv("B200_GROUP_WARPS_PER_SM", str(16 // 4 if world == 3 else 16 // world))
    m, n = 30000, 26000
    a, b = _pair(m, n, 41)
    g = _group(b200, world, m, n, chunk, _kern(b200, kernel))
    g.set_sequences(a, b)
    o = None
    for rep, prune in enumerate((False, True, False)):
        if prune and kernel == "s32":
            continue
        r = g.align_partition(want_last_row=True, want_last_column=True, want_special_rows=True, special_row_interval=1000,
                              mgpu=False, chunk_cols=chunk, prune=prune)
        ids = sorted(i for i in r["rows"] if i != m)
        if o is None:
            o = O.full_matrix(a, b, O.SW, row_ids=[i - 1 for i in ids] + [m - 1])
        assert r["best"] == o["best"], (rep, prune)
        assert r["scores"] == [o["best"]]
        if not prune:
            assert r["cells"] == m * n
            for i in ids + [m]:
                assert np.array_equal(r["rows"][i], o["rows"][i - 1]), f"row {i}"
            assert np.array_equal(r["last_column"], o["last_col"])
            per = g.rank_results()
            assert sum(p["cells"] for p in per) == m * n
            assert max(p["cells"] for p in per) <= 1.35 * m * n / world or chunk == -1 or world == 3
    g.close()


def test_group_on_one_device_pruned_balance(b200, short_watchdog, monkeypatch):
    """Round-robin chunks spread the cells that survive pruning over the ranks (what static slices cannot do)."""
    world, chunk = 4, 1024
    monkeypatch.setenv("B200_GROUP_WARPS_PER_SM", "4")
    m, n = 90000, 90000
    a, b = _pair(m, n, 43, 0.02, 0.98)
    o = O.full_matrix(a, b, O.SW, want_last_col=False)
    g = _group(b200, world, m, n, chunk, b200.KERNEL_S16X2)
    g.set_sequences(a, b)
    r = g.align_partition(use_callbacks=False, chunk_cols=chunk, prune=True)
    assert r["best"] == o["best"]
    per = [p["cells"] for p in g.rank_results()]
    assert sum(per) == r["cells"] < 0.8 * m * n
    assert max(per) <= 1.15 * sum(per) / world, per
    g.close()


def test_group_nw_global_with_special_rows(b200, short_watchdog, monkeypatch):
    monkeypatch.setenv("B200_GROUP_WARPS_PER_SM", "8")
    m, n = 20000, 21000
    a, b = _pair(m, n, 13, 0.0, 1.0)
    g = _group(b200, 2, m, n, 3000, b200.KERNEL_AUTO)
    g.set_sequences(a, b)
    r = g.align_partition(recurrence=b200.NEEDLEMAN_WUNSCH, first_row_init=b200.INIT_GAPS, first_col_init=b200.INIT_GAPS,
                          want_last_row=True, want_last_column=True, want_best_score=False, want_special_rows=True,
                          special_row_interval=1000, chunk_cols=3000)
    ids = sorted(i for i in r["rows"] if i != m)
    assert ids == [8192, 16384]
    o = O.full_matrix(a, b, O.NW, first_row_type=O.INIT_GAPS, first_col_type=O.INIT_GAPS, row_ids=[i - 1 for i in ids] + [m - 1])
    for i in ids + [m]:
        assert np.array_equal(r["rows"][i], o["rows"][i - 1]), f"row {i}"
    assert np.array_equal(r["last_column"], o["last_col"])
    g.close()
