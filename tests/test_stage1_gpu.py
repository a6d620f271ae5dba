"""GPU parity of the whole-partition path (b200_align_partition) against the plain-C oracle.
Bit-exact: best score + coordinates, special rows, last row, last column."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import synth  # noqa: E402

pytestmark = pytest.mark.gpu


def _pair(m, n, seed, frac=0.6):
    a0 = int(m * 0.2)
    a1 = int(m * (0.2 + frac))
    return synth.make_pair(m, n, [(a0, a1)], 0.05, 0.01, 0.01, 0, seed)


@pytest.mark.parametrize("kernel", ["s32", "s16x2"])
@pytest.mark.parametrize("m,n,seed", [(700, 900, 1), (1, 1, 2), (33, 5, 3), (512, 512, 4), (513, 31, 5), (2049, 1500, 6), (5000, 7000, 7)])
def test_sw_best_and_borders(b200, kernel, m, n, seed):
    a, b = _pair(m, n, seed)
    al = b200.Aligner(kernel=b200.KERNEL_S32 if kernel == "s32" else b200.KERNEL_S16X2)
    al.set_sequences(a, b)
    r = al.align_partition(want_last_row=True, want_last_column=True)
    o = O.full_matrix(a, b, O.SW, row_ids=[m - 1])
    assert r["best"] == o["best"]
    assert r["cells"] == m * n
    lr = r["rows"][m]
    assert np.array_equal(lr, o["rows"][m - 1])
    assert np.array_equal(r["last_column"], o["last_col"])
    al.close()


@pytest.mark.parametrize("kernel", ["s32", "s16x2"])
@pytest.mark.parametrize("m,n", [(9000, 3000), (20000, 1111)])
def test_special_rows(b200, kernel, m, n):
    a, b = _pair(m, n, 11)
    al = b200.Aligner(kernel=b200.KERNEL_S32 if kernel == "s32" else b200.KERNEL_S16X2)
    al.set_sequences(a, b)
    r = al.align_partition(want_special_rows=True, special_row_interval=1000, want_last_row=True)
    ids = sorted(i for i in r["rows"] if i != m)
    assert ids == list(range(8192, m, 8192))           # 8192-row floor of AbstractDiagonalAligner.cpp:35,466-478
    o = O.full_matrix(a, b, O.SW, row_ids=[i - 1 for i in ids] + [m - 1])
    for i in ids + [m]:
        assert np.array_equal(r["rows"][i], o["rows"][i - 1]), f"special row {i}"
    assert r["best"] == o["best"]
    al.close()


@pytest.mark.parametrize("kernel", ["s32", "s16x2"])
@pytest.mark.parametrize("rt,ct", [(O.INIT_GAPS, O.INIT_GAPS), (O.INIT_GAPS_OPENED, O.INIT_GAPS), (O.INIT_GAPS, O.INIT_ZEROES)])
def test_nw_global(b200, kernel, rt, ct):
    m, n = 3001, 2500
    a, b = _pair(m, n, 21, frac=0.75)
    al = b200.Aligner(kernel=b200.KERNEL_S32 if kernel == "s32" else b200.KERNEL_S16X2)
    al.set_sequences(a, b)
    r = al.align_partition(recurrence=b200.NEEDLEMAN_WUNSCH, first_row_init=rt, first_col_init=ct,
                           want_last_row=True, want_last_column=True, want_best_score=False)
    o = O.full_matrix(a, b, O.NW, first_row_type=rt, first_col_type=ct, row_ids=[m - 1])
    assert np.array_equal(r["rows"][m], o["rows"][m - 1])
    assert np.array_equal(r["last_column"], o["last_col"])
    al.close()


def test_custom_borders_subpartition(b200):
    """A partition in the middle of the sequences with caller-supplied first row / column (stage 2/3 shape)."""
    a, b = _pair(4000, 4000, 31)
    i0, j0, i1, j1 = 700, 300, 3100, 3333
    rng = np.random.default_rng(5)
    fr = np.zeros(j1 - j0 + 1, O.CELL); fc = np.zeros(i1 - i0 + 1, O.CELL)
    fr["h"] = -np.cumsum(rng.integers(0, 4, fr.size)); fr["x"] = fr["h"] - rng.integers(1, 9, fr.size)
    fc["h"] = -np.cumsum(rng.integers(0, 4, fc.size)); fc["x"] = fc["h"] - rng.integers(1, 9, fc.size)
    fc[0] = fr[0]
    for kernel in (b200.KERNEL_S32, b200.KERNEL_S16X2):
        al = b200.Aligner(kernel=kernel)
        al.set_sequences(a, b)
        r = al.align_partition(i0, j0, i1, j1, recurrence=b200.NEEDLEMAN_WUNSCH, first_row_init=b200.INIT_CUSTOM,
                               first_col_init=b200.INIT_CUSTOM, first_row=fr, first_col=fc, want_last_row=True,
                               want_last_column=True, want_best_score=True)
        o = O.full_matrix(a[i0:i1], b[j0:j1], O.NW, first_row=fr, first_col=fc, row_ids=[i1 - i0 - 1])
        assert np.array_equal(r["rows"][i1][1:], o["rows"][i1 - i0 - 1][1:])
        assert np.array_equal(r["last_column"][1:], o["last_col"][1:])
        assert r["best"] == (o["best"][0], o["best"][1] + i0, o["best"][2] + j0)
        al.close()


@pytest.mark.parametrize("m,n,hom", [(60000, 50000, (5000, 45000)), (30000, 90000, (1000, 29000)), (150000, 150000, (20000, 130000)), (20000, 20000, (0, 0))])
def test_pruning_keeps_best_exact(b200, m, n, hom):
    """On-device block pruning (SW): the best cell must stay bit-exact, cells are skipped, and every published
    bottom-row cell is a lower bound of the exact one (skipped cells are published as H = 0)."""
    a, b = synth.make_pair(m, n, [hom], 0.05, 0.01, 0.01, 0, 5)
    al = b200.Aligner(kernel=b200.KERNEL_S16X2)
    al.set_sequences(a, b)
    exact = al.align_partition(want_last_row=True, want_last_column=True, prune=False)
    r = al.align_partition(want_last_row=True, want_last_column=True, prune=True)
    assert r["best"] == exact["best"]
    assert r["cells"] <= exact["cells"]
    if hom[1] - hom[0] > 20000:
        assert r["cells"] < 0.9 * exact["cells"], "a long planted alignment must let the strips skip blocks"
    assert np.all(r["rows"][m]["h"][1:] <= exact["rows"][m]["h"][1:])
    assert np.all(r["last_column"]["h"][1:] <= exact["last_column"]["h"][1:])
    assert np.all(r["rows"][m]["h"][1:] >= 0)
    if m * n <= 4_000_000_000:
        o = O.full_matrix(a, b, O.SW, want_last_col=False)
        assert r["best"] == o["best"]
    al.close()


def test_pruning_keeps_best_exact_with_million_scores(b200):
    """Regression (found on the 23M x 25M BASELINE pair): in the lower part of a large, highly similar matrix the
    surviving band starts at cells worth millions.  A compute segment restarted there with its s16 frame anchored at
    zero saturated for tmax/32767 blocks, the strips below inherited the under-estimates and the damaged front crept
    into the alignment path: the pruned run reported a best cell ~5 % short of the exact one.  3M x 3M at 99 % identity
    reproduces it in seconds; the exact answer is the same kernel with pruning off (itself pinned to the oracle above)."""
    m = n = 3_000_000
    a, b = synth.make_pair(m, n, [(0, m)], 0.01, 0.0, 0.0, 0, 41)
    al = b200.Aligner(kernel=b200.KERNEL_S16X2)
    al.set_sequences(a, b)
    exact = al.align_partition(prune=False, use_callbacks=False)
    r = al.align_partition(prune=True, use_callbacks=False)
    assert exact["best"][0] > 2_500_000
    assert r["best"] == exact["best"]
    assert r["cells"] < 0.8 * exact["cells"]
    al.close()


def test_mixed_alphabet_matches_byte_compare(b200):
    """Real FASTA files carry N runs and IUPAC codes; the reference compares raw bytes (N == N matches).  Strips whose
    rows hold such bytes take the int32 code path inside the same launch, the others stay on the packed kernel."""
    m, n = 9000, 7000
    a, b = synth.make_pair(m, n, [(1000, 8000)], 0.05, 0.01, 0.01, 0, 17)
    a = a.copy(); b = b.copy()
    a[2000:2300] = ord("N"); b[1700:2100] = ord("N")          # overlapping N runs: N-N cells score +1
    a[5000] = ord("R"); a[5001] = ord("n"); b[4000:4005] = np.frombuffer(b"RYKMN", dtype=np.uint8)
    al = b200.Aligner()                                        # KERNEL_AUTO
    al.set_sequences(a, b)
    r = al.align_partition(want_last_row=True, want_last_column=True, want_special_rows=True, special_row_interval=1000)
    assert r["kernel_used"] == b200.KERNEL_S16X2               # mixed launch: most strips are still packed
    o = O.full_matrix(a, b, O.SW, row_ids=[8191, m - 1])
    assert r["best"] == o["best"]
    assert np.array_equal(r["rows"][8192], o["rows"][8191])
    assert np.array_equal(r["rows"][m], o["rows"][m - 1])
    assert np.array_equal(r["last_column"], o["last_col"])
    rp = al.align_partition(prune=True)
    assert rp["best"] == o["best"]
    al.close()


@pytest.fixture(scope="module")
def chain_case():
    m, n = 70000, 40000                                       # 69 strips of 1024 rows, 1250 blocks each
    a, b = synth.make_pair(m, n, [(4000, 60000)], 0.05, 0.01, 0.01, 0, 31)
    ids = list(range(8192, m, 8192)) + [m]
    return a, b, ids, O.full_matrix(a, b, O.SW, row_ids=[i - 1 for i in ids])   # the oracle runs once for all variants


@pytest.mark.parametrize("opt", [0, 1, 2, 3, 7, 15, 27, 31, 32 + 27, 64 + 27, 96 + 27, 96 + 31])
def test_chain_protocol_variants_are_exact(b200, opt, chain_case, monkeypatch):
    """Every StripOpt combination (csrc/strip_common.cuh: fence choice, cached progress, 128-column releases, best
    exchange every 4th block, 128-column skips, look-ahead waits, deferred releases) only changes how often the strip chain synchronises: results stay
    bit-exact without pruning, and the best cell stays exact with pruning."""
    a, b, ids, o = chain_case
    monkeypatch.setenv("B200_OPT", str(opt))
    al = b200.Aligner(kernel=b200.KERNEL_S16X2)
    al.set_sequences(a, b)
    r = al.align_partition(want_last_row=True, want_last_column=True, want_special_rows=True, special_row_interval=1000)
    assert sorted(r["rows"]) == ids
    assert r["best"] == o["best"]
    for i in ids:
        assert np.array_equal(r["rows"][i], o["rows"][i - 1]), f"row {i}"
    assert np.array_equal(r["last_column"], o["last_col"])
    rp = al.align_partition(prune=True)
    assert rp["best"] == o["best"]
    assert rp["cells"] < r["cells"]
    al.close()


@pytest.mark.parametrize("m,n", [(5000, 7001), (33, 15), (20000, 16)])
def test_packed_and_byte_sequences_agree(b200, monkeypatch, m, n):
    """Pure A/C/G/T inputs live 2-bit packed in HBM (16 bases per word) and the packed kernel reads those words; with
    B200_NO_PACK the same kernel reads the byte arrays.  Both must match the oracle (odd lengths: partial last words)."""
    a, b = _pair(m, n, 9)
    o = O.full_matrix(a, b, O.SW, row_ids=[m - 1])
    for no_pack in ("", "1"):
        if no_pack:
            monkeypatch.setenv("B200_NO_PACK", "1")
        else:
            monkeypatch.delenv("B200_NO_PACK", raising=False)
        al = b200.Aligner(kernel=b200.KERNEL_S16X2)
        al.set_sequences(a, b)
        r = al.align_partition(want_last_row=True, want_last_column=True)
        assert r["best"] == o["best"]
        assert np.array_equal(r["rows"][m], o["rows"][m - 1])
        assert np.array_equal(r["last_column"], o["last_col"])
        r = al.align_partition(17 % m, 3 % n, m, n, want_last_row=True, use_callbacks=True, want_best_score=True)     # unaligned sub-partition
        o2 = O.full_matrix(a[17 % m:], b[3 % n:], O.SW, row_ids=[m - 17 % m - 1])
        assert r["best"] == (o2["best"][0], o2["best"][1] + 17 % m, o2["best"][2] + 3 % n)
        al.close()


def test_nw_border_with_minus_inf_routes_to_int32(b200):
    """An NW partition whose first column / first row carries -INF in H (borders handed over by a pruned neighbour): the
    16-bit frame of the packed kernel cannot drift like the reference's plain int32 arithmetic, so the engine must run
    such a partition on the int32 kernel by itself -- also when the handle asks for the packed one."""
    a, b = _pair(3000, 2600, 41)
    i0, j0, i1, j1 = 100, 200, 2900, 2500
    rng = np.random.default_rng(7)
    fr = np.zeros(j1 - j0 + 1, O.CELL); fc = np.zeros(i1 - i0 + 1, O.CELL)
    fr["h"] = -np.cumsum(rng.integers(0, 4, fr.size)); fr["x"] = fr["h"] - rng.integers(1, 9, fr.size)
    fc["h"] = -np.cumsum(rng.integers(0, 4, fc.size)); fc["x"] = fc["h"] - rng.integers(1, 9, fc.size)
    fc[0] = fr[0]
    fc["h"][700:1500] = -O.INF; fc["x"][700:1500] = -O.INF
    fr["h"][1000:1300] = -O.INF; fr["x"][1000:1300] = -O.INF
    o = O.full_matrix(a[i0:i1], b[j0:j1], O.NW, first_row=fr, first_row_type=O.INIT_CUSTOM, first_col=fc,
                      first_col_type=O.INIT_CUSTOM, row_ids=[i1 - i0 - 1])
    for kernel in (b200.KERNEL_S16X2, b200.KERNEL_AUTO):
        al = b200.Aligner(kernel=kernel)
        al.set_sequences(a, b)
        r = al.align_partition(i0, j0, i1, j1, recurrence=b200.NEEDLEMAN_WUNSCH, first_row_init=b200.INIT_CUSTOM,
                               first_col_init=b200.INIT_CUSTOM, first_row=fr, first_col=fc, want_last_row=True,
                               want_last_column=True, want_best_score=False)
        assert r["kernel_used"] == b200.KERNEL_S32
        assert np.array_equal(r["rows"][i1], o["rows"][i1 - i0 - 1])
        assert np.array_equal(r["last_column"], o["last_col"])
        al.close()
