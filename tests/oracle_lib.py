"""ctypes access to the plain-C oracle (oracle/_ref/libgotoh_oracle.so) and to the reference-built binaries.
TEST INFRASTRUCTURE: imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
CELL = np.dtype([("h", "<i4"), ("x", "<i4")])
INF = 999999999
NW, SW = 0, 1
INIT_ZEROES, INIT_GAPS, INIT_CUSTOM, INIT_GAPS_OPENED = 0, 1, 2, 3


class GoScore(C.Structure):
    _fields_ = [("score", C.c_int), ("i", C.c_int), ("j", C.c_int)]


class GoMatch(C.Structure):
    _fields_ = [("found", C.c_int), ("k", C.c_int), ("score", C.c_int), ("type", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(REF_DIR, "libgotoh_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
        _lib = C.CDLL(path)
        _lib.go_full_matrix.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.POINTER(GoScore)]
        _lib.go_init_cells.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.go_init_cells.restype = None
        _lib.go_match_column.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.go_match_column.restype = GoMatch
    return _lib


def init_cells(n, kind, start=0):
    out = np.zeros(n, CELL)
    lib().go_init_cells(out.ctypes.data, n, kind, start)
    return out


def full_matrix(s0, s1, recurrence=SW, first_row=None, first_row_type=INIT_ZEROES, first_col=None,
                first_col_type=INIT_ZEROES, row_ids=(), want_last_col=True):
    """Returns dict(best=(score,i,j), rows={row_id: CELL[n+1]}, last_col=CELL[m+1])."""
    a = np.ascontiguousarray(s0, dtype=np.uint8)
    b = np.ascontiguousarray(s1, dtype=np.uint8)
    m, n = a.size, b.size
    ids = np.ascontiguousarray(sorted(row_ids), dtype=np.int32)
    rows = np.zeros((ids.size, n + 1), CELL)
    last = np.zeros(m + 1, CELL) if want_last_col else None
    best = GoScore()
    fr = np.ascontiguousarray(first_row, dtype=CELL) if first_row is not None else None
    fc = np.ascontiguousarray(first_col, dtype=CELL) if first_col is not None else None
    rc = lib().go_full_matrix(a.ctypes.data, m, b.ctypes.data, n, recurrence,
                              fr.ctypes.data if fr is not None else None, first_row_type,
                              fc.ctypes.data if fc is not None else None, first_col_type,
                              ids.ctypes.data if ids.size else None, ids.size, rows.ctypes.data if ids.size else None,
                              last.ctypes.data if last is not None else None, C.byref(best))
    assert rc == 0
    return dict(best=(best.score, best.i, best.j), rows={int(i): rows[k] for k, i in enumerate(ids)}, last_col=last)


def match_column(buffer, base, goal, gap_open=3):
    bu = np.ascontiguousarray(buffer, dtype=CELL)
    ba = np.ascontiguousarray(base, dtype=CELL)
    r = lib().go_match_column(bu.ctypes.data, ba.ctypes.data, bu.size, goal, gap_open)
    return dict(found=bool(r.found), k=r.k, score=r.score, type=r.type)


def have_ref_binaries():
    return os.path.exists(os.path.join(REF_DIR, "oracle_cpu")) and os.path.exists(os.path.join(REF_DIR, "oracle_cpu_block"))


def run_ref(binary, fasta_a, fasta_b, workdir, extra=(), env=None):
    """Run oracle/_ref/<binary> (the reference's own CPU path) and return the work directory."""
    exe = os.path.join(REF_DIR, binary)
    cmd = [exe, f"--work-dir={workdir}", "--clear", "--verbose=0", *extra, fasta_a, fasta_b]
    e = dict(os.environ)
    if env:
        e.update(env)
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=e)
    return workdir


def read_crosspoints(path):
    pts = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line in ("START", "END"):
                continue
            t, i, j, s = (int(x) for x in line.split(","))
            pts.append((t, i, j, s))
    return pts


XPOINT = np.dtype([("i", "<i4"), ("j", "<i4"), ("type", "<i4"), ("score", "<i4")])     # == crosspoint_t


def stage4_round(s0, s1, points, max_part=16):
    """One reduce_partitions + merge_partitions round of the reference's stage 4 (C restatement). Returns (points, changed)."""
    L = lib()
    L.go_stage4_round.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    a = np.ascontiguousarray(s0, dtype=np.uint8); b = np.ascontiguousarray(s1, dtype=np.uint8)
    pts = np.ascontiguousarray(points, dtype=XPOINT)
    out = np.zeros(2 * pts.size + 2, XPOINT)
    ch = C.c_int()
    n = L.go_stage4_round(a.ctypes.data, b.ctypes.data, pts.ctypes.data, pts.size, max_part, out.ctypes.data, C.byref(ch))
    assert n > 0, f"stage-4 oracle failed with {n}"
    return out[:n].copy(), bool(ch.value)


def largest_partition(points):
    L = lib()
    L.go_largest_partition.argtypes = [C.c_void_p, C.c_int]
    pts = np.ascontiguousarray(points, dtype=XPOINT)
    return L.go_largest_partition(pts.ctypes.data, pts.size)


def stage4(s0, s1, points, max_part=16):
    pts = np.ascontiguousarray(points, dtype=XPOINT)
    while largest_partition(pts) > max_part:
        pts, changed = stage4_round(s0, s1, pts, max_part)
        if not changed:
            break
    return pts


def golden_points(entries):
    """[(type, i, j, score), ...] as stored in tests/golden -> XPOINT array."""
    return np.array([(i, j, t, s) for (t, i, j, s) in entries], dtype=XPOINT)


class GoS5Stats(C.Structure):
    _fields_ = [("score", C.c_int), ("matches", C.c_int), ("mismatches", C.c_int), ("gap_open", C.c_int), ("gap_ext", C.c_int)]


def stage5(s0, s1, points):
    """The reference's stage-5 traceback (C restatement go_stage5) over the partitions between consecutive crosspoints.
    Returns (ops, op_off, op_len, stats): partition k (between points k-1 and k) owns ops[op_off[k] : op_off[k] + op_len[k]],
    one byte per step from its bottom-right corner (0 diagonal, 1 vertical = gap in seq1, 2 horizontal = gap in seq0)."""
    L = lib()
    L.go_stage5.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(GoS5Stats)]
    a = np.ascontiguousarray(s0, dtype=np.uint8); b = np.ascontiguousarray(s1, dtype=np.uint8)
    pts = np.ascontiguousarray(points, dtype=XPOINT)
    cap = int(pts["i"][-1] - pts["i"][0]) + int(pts["j"][-1] - pts["j"][0])
    ops = np.full(cap + 1, 255, np.uint8)
    op_len = np.zeros(pts.size, np.int32)
    st = GoS5Stats()
    rc = L.go_stage5(a.ctypes.data, b.ctypes.data, pts.ctypes.data, pts.size, ops.ctypes.data, op_len.ctypes.data, C.byref(st))
    assert rc == 0, f"stage-5 oracle failed with {rc}"
    return ops[:cap], stage5_offsets(pts), op_len, {k: getattr(st, k) for k, _ in GoS5Stats._fields_}


def stage5_offsets(points):
    """op_off[k] = (i[k-1] - i[0]) + (j[k-1] - j[0]): every partition owns as many slots as its longest possible walk."""
    pts = np.ascontiguousarray(points, dtype=XPOINT)
    off = np.zeros(pts.size, np.int64)
    off[1:] = (pts["i"][:-1].astype(np.int64) - int(pts["i"][0])) + (pts["j"][:-1].astype(np.int64) - int(pts["j"][0]))
    return off


def stage5_gaps(points, ops, op_off, op_len):
    """What stage5() hands to Alignment (dot(), C/stage5/sw_stage5.cpp:70-84, forward sequences; Alignment::addGap run-length
    merge, C/common/biology/Alignment.cpp, then finalize()'s sort by position): (gaps0, gaps1) as lists of [pos, len].
    A gap run cut by a crosspoint yields two entries with the same position (the partitions are visited top-down, each one
    walked bottom-up, so the two halves are not consecutive calls); finalize() orders such twins by the whim of std::sort,
    so compare the lists sorted by (pos, len)."""
    pts = np.ascontiguousarray(points, dtype=XPOINT)
    gaps = ([], [])
    for k in range(1, pts.size):
        i, j = int(pts["i"][k] - pts["i"][k - 1]), int(pts["j"][k] - pts["j"][k - 1])
        i0, j0 = int(pts["i"][k - 1]), int(pts["j"][k - 1])
        for d in ops[int(op_off[k]):int(op_off[k]) + int(op_len[k])]:
            if d == 0:
                i -= 1; j -= 1
                continue
            seq, pos = (1, j0 + j + 1) if d == 1 else (0, i0 + i + 1)
            g = gaps[seq]
            if g and g[-1][0] == pos:
                g[-1][1] += 1
            else:
                g.append([pos, 1])
            if d == 1:
                i -= 1
            else:
                j -= 1
        assert i == 0 and j == 0, f"partition {k}: the walk stopped at ({i},{j})"
    return sorted(gaps[0]), sorted(gaps[1])
