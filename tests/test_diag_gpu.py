"""GPU parity of the diag primitives (one C-ABI call per CUDAligner virtual) against the plain-C oracle.
The loop below is what AbstractDiagonalAligner::processNextIteration does (AbstractDiagonalAligner.cpp:110-159):
load the first-column chunk, process one external diagonal, then read special/last rows, the last-column chunk
and the block scores."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import synth  # noqa: E402

pytestmark = pytest.mark.gpu


def split_evenly(j0, j1, blocks):          # AlignerUtils::splitBlocksEvenly (utils/AlignerUtils.cpp:38-45)
    n = j1 - j0
    mod = n % blocks
    pos = [j0 + (n // blocks) * i + (i if i < mod else mod) for i in range(blocks)]
    return pos + [j1]


def run_diag(b200, al, a, b, i0, j0, i1, j1, rec, fr, fc, frt, fct, B, bh, window=None):
    m, n = i1 - i0, j1 - j0
    part = b200.Partition(i0=i0, j0=j0, i1=i1, j1=j1, recurrence=rec, first_row_init=frt, first_col_init=fct)
    split = split_evenly(j0, j1, B)
    al.diag_set_first_row(fr[1:], j0)               # prepareIterations loads the first row before initializeDiagonals
    al.diag_begin(part, split, bh)
    gh = m // bh + 1
    last_col = [np.array([(fr["h"][-1], -O.INF)], dtype=O.CELL)]
    last_row = np.zeros(n, O.CELL)
    best = (-O.INF, -1, -1)
    for d in range(gh + B):
        if fct != b200.INIT_ZEROES:
            pos = d * bh
            if pos < m:
                ln = min(bh, m - pos)
                chunk = np.zeros(bh + 1, O.CELL)
                chunk[0] = fc[pos]                  # diagonal cell = first-column tail
                chunk[1:1 + ln] = fc[pos + 1:pos + 1 + ln]
                chunk[1 + ln:] = (-O.INF, -O.INF)
                al.diag_set_first_column(chunk, pos, ln)
        wl, wr = window(d) if window else (0, B)
        al.diag_process(d, wl, wr)
        # last row: block (bx-1, gh-1) finished in this call when bx-1 + gh-1 == d-1
        for bx in range(B):
            by = d - 1 - bx
            if by >= 0 and by * bh < m and (by + 1) * bh >= m:
                last_row[split[bx] - j0:split[bx + 1] - j0] = al.diag_get_row(split[bx], split[bx + 1] - split[bx])
        i = (d - B) * bh
        if 0 <= i < m:
            ln = min(bh, m - i)
            last_col.append(al.diag_get_last_column(i, ln))
        sc = al.diag_get_block_scores(B)
        for s in sc:
            if s["i"] >= 0:
                cand = (int(s["score"]), int(s["i"]), int(s["j"]))
                if cand[0] > best[0] or (cand[0] == best[0] and (cand[1], cand[2]) < (best[1], best[2])):
                    best = cand
    al.diag_end()
    return np.concatenate(last_col), last_row, best


@pytest.mark.parametrize("kernel", ["s32", "s16x2"])
@pytest.mark.parametrize("rec,frt,fct", [("sw", 0, 0), ("nw", 1, 1), ("nw", 3, 1), ("nw", 1, 3)])
@pytest.mark.parametrize("m,n,B,bh", [(3000, 2700, 5, 512), (5000, 1300, 2, 512), (700, 100, 1, 400), (1537, 4096, 8, 512)])
def test_diag_matches_oracle(b200, kernel, rec, frt, fct, m, n, B, bh):
    a, b = synth.make_pair(m + 100, n + 50, [(300, m - 200)], 0.05, 0.02, 0.02, 0, 5)
    i0, j0 = 40, 20
    i1, j1 = i0 + m, j0 + n
    recurrence = b200.SMITH_WATERMAN if rec == "sw" else b200.NEEDLEMAN_WUNSCH
    fr = O.init_cells(n + 1, frt)
    fc = O.init_cells(m + 1, fct)
    al = b200.Aligner(kernel=b200.KERNEL_S32 if kernel == "s32" else b200.KERNEL_S16X2)
    al.set_sequences(a, b)
    lc, lr, best = run_diag(b200, al, a, b, i0, j0, i1, j1, recurrence, fr, fc, frt, fct, B, bh)
    o = O.full_matrix(a[i0:i1], b[j0:j1], O.SW if rec == "sw" else O.NW, first_row=fr, first_col=fc, row_ids=[m - 1])
    assert np.array_equal(lc, o["last_col"]), np.nonzero((lc["h"] != o["last_col"]["h"]) | (lc["x"] != o["last_col"]["x"]))[0][:10]
    assert np.array_equal(lr, o["rows"][m - 1][1:])
    assert best == (o["best"][0], o["best"][1] + i0, o["best"][2] + j0)
    al.close()
