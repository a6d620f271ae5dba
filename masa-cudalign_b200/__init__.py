"""ctypes binding of libb200align.so (the C ABI in include/b200align.h) for tests, bench.py and smoke().

This is NOT the product's host layer -- that is the C++ adapter in host/ (B200Aligner, the drop-in for the
reference's CUDAligner class, R/src/CUDAligner.cpp).  Python is only used to drive the library from pytest and
from the benchmark.  There is no CPU fallback here: if the shared library is missing or no GPU is present the
calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(_HERE, "libb200align.so")   # B200_LIB: development builds of the same library

INF = 999999999
NEEDLEMAN_WUNSCH, SMITH_WATERMAN = 0, 1                          # C/libmasa/IManager.hpp:31-33
INIT_ZEROES, INIT_GAPS, INIT_CUSTOM, INIT_GAPS_OPENED = 0, 1, 2, 3    # C/libmasa/IManager.hpp:38-47
KERNEL_AUTO, KERNEL_S32, KERNEL_S16X2 = 0, 1, 2

CELL = np.dtype([("h", "<i4"), ("x", "<i4")])
XPOINT = np.dtype([("i", "<i4"), ("j", "<i4"), ("type", "<i4"), ("score", "<i4")])      # == crosspoint_t


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("kernel", C.c_int), ("warps_per_sm", C.c_int), ("reserved", C.c_int * 5)]


class Partition(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "i0", "j0", "i1", "j1", "recurrence", "first_row_init", "first_col_init", "special_row_interval",
        "block_height", "want_special_rows", "want_last_row", "want_last_column", "want_best_score", "prune",
        "super_i1", "super_j1")] + [("reserved", C.c_int * 4)]


class Score(C.Structure):
    _fields_ = [("i", C.c_int), ("j", C.c_int), ("score", C.c_int)]


class Match(C.Structure):
    _fields_ = [("found", C.c_int), ("k", C.c_int), ("score", C.c_int), ("type", C.c_int)]


class Result(C.Structure):
    _fields_ = [("best", Score), ("cells", C.c_longlong), ("cells_total", C.c_longlong), ("device_ms", C.c_double),
                ("strips", C.c_int), ("kernel_launches", C.c_int), ("kernel_used", C.c_int), ("reserved", C.c_int * 5)]


class ChainInfo(C.Structure):
    _fields_ = [("chunks", C.c_int), ("chunk_cols", C.c_int), ("chunks_per_gpu", C.c_int), ("reserved0", C.c_int),
                ("max_strips", C.c_longlong), ("max_jobs", C.c_longlong)]


class S5Stats(C.Structure):
    _fields_ = [("score", C.c_int), ("matches", C.c_int), ("mismatches", C.c_int), ("gap_open", C.c_int), ("gap_ext", C.c_int)]


RECV_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int)
DISP_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p, C.c_int)
SCORE_FN = C.CFUNCTYPE(None, C.c_void_p, Score)
CONT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)


class Callbacks(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("receive_first_row", RECV_FN), ("receive_first_column", RECV_FN),
                ("dispatch_row", DISP_FN), ("dispatch_column", DISP_FN), ("dispatch_score", SCORE_FN),
                ("must_continue", CONT_FN)]


EXPORTS = [
    "b200_create", "b200_destroy", "b200_last_error", "b200_device_count", "b200_set_sequences",
    "b200_unset_sequences", "b200_align_partition", "b200_diag_begin", "b200_diag_set_first_row",
    "b200_diag_set_first_column", "b200_diag_process", "b200_diag_get_row", "b200_diag_get_last_column",
    "b200_diag_get_block_scores", "b200_diag_clear_pruned", "b200_diag_end", "b200_match_last_column",
    "b200_processed_cells", "b200_kernel_launches", "b200_chain_plan", "b200_mgpu_export", "b200_mgpu_connect", "b200_mgpu_disconnect",
    "b200_last_chain_result", "b200_group_create", "b200_group_destroy", "b200_group_last_error", "b200_group_size", "b200_group_handle",
    "b200_group_set_sequences", "b200_group_align_partition", "b200_group_rank_result",
    "b200_special_row_ids", "b200_stage4_round", "b200_stage4", "b200_stage5",
]

_lib = None


def load_library(path=None):
    """Load libb200align.so (raises OSError when it has not been built: there is no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    lib = C.CDLL(path or LIB_PATH)
    if hasattr(lib, "b200_emu_is_emulation") and os.environ.get("B200_TEST_EMULATION") != "1":
        # tests/emu/ builds the same sources for the host CPU so that the test suite can exercise the device code without a
        # GPU.  That build is test infrastructure: only a test run that says so may bind it.
        raise OSError(f"{path or LIB_PATH} is the SIMT emulation build of the test suite, not libb200align.so: this package has "
                      "no CPU path (tests set B200_TEST_EMULATION=1 to bind it)")
    lib.b200_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    lib.b200_destroy.argtypes = [C.c_void_p]
    lib.b200_destroy.restype = None
    lib.b200_last_error.argtypes = [C.c_void_p]
    lib.b200_last_error.restype = C.c_char_p
    lib.b200_set_sequences.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.b200_unset_sequences.argtypes = [C.c_void_p]
    lib.b200_align_partition.argtypes = [C.c_void_p, C.POINTER(Partition), C.POINTER(Callbacks), C.POINTER(Result)]
    lib.b200_diag_begin.argtypes = [C.c_void_p, C.POINTER(Partition), C.c_int, C.c_void_p, C.c_int]
    lib.b200_diag_set_first_row.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.b200_diag_set_first_column.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.b200_diag_process.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.b200_diag_get_row.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.b200_diag_get_last_column.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.b200_diag_get_block_scores.argtypes = [C.c_void_p, C.c_void_p]
    lib.b200_diag_clear_pruned.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.b200_diag_end.argtypes = [C.c_void_p]
    lib.b200_match_last_column.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(Match)]
    lib.b200_processed_cells.argtypes = [C.c_void_p]
    lib.b200_processed_cells.restype = C.c_longlong
    lib.b200_kernel_launches.argtypes = [C.c_void_p]
    lib.b200_kernel_launches.restype = C.c_longlong
    lib.b200_chain_plan.argtypes = [C.POINTER(Partition), C.c_int, C.POINTER(ChainInfo)]
    lib.b200_mgpu_export.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p]
    lib.b200_last_chain_result.argtypes = [C.c_void_p, C.POINTER(Result)]
    lib.b200_group_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(Config), C.c_longlong, C.c_longlong, C.POINTER(C.c_void_p)]
    lib.b200_group_destroy.argtypes = [C.c_void_p]
    lib.b200_group_destroy.restype = None
    lib.b200_group_last_error.argtypes = [C.c_void_p]
    lib.b200_group_last_error.restype = C.c_char_p
    lib.b200_group_size.argtypes = [C.c_void_p]
    lib.b200_group_handle.argtypes = [C.c_void_p, C.c_int]
    lib.b200_group_handle.restype = C.c_void_p
    lib.b200_group_set_sequences.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.b200_group_align_partition.argtypes = [C.c_void_p, C.POINTER(Partition), C.POINTER(Callbacks), C.POINTER(Result)]
    lib.b200_group_rank_result.argtypes = [C.c_void_p, C.c_int, C.POINTER(Result)]
    lib.b200_mgpu_connect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.b200_mgpu_disconnect.argtypes = [C.c_void_p]
    lib.b200_special_row_ids.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.b200_stage4_round.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.b200_stage4.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.b200_stage5.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.POINTER(S5Stats)]
    if path is None:
        _lib = lib
    return lib


def special_row_ids(height, block_height, interval):
    """Row ids (rows above the special row, relative to the partition) of the reference's flush policy."""
    lib = load_library()
    n = lib.b200_special_row_ids(height, block_height, interval, None, 0)
    out = (C.c_int * max(n, 1))()
    lib.b200_special_row_ids(height, block_height, interval, out, n)
    return [out[k] for k in range(n)]


def column_slice(n, rank, world):
    """Columns [j0, j1) of `rank` when the chain runs with one contiguous slice per GPU (chunk_cols < 0): equal
    weights, the integer arithmetic of the reference's --split/--fork (C/libmasa/libmasa.cpp:632-635)."""
    return n * rank // world, n * (rank + 1) // world


def chain_plan(m, n, world, chunk_cols=0):
    """Column chunks of a chained m x n partition (host-side planning, no GPU): dict(chunks, chunk_cols,
    chunks_per_gpu, max_strips, max_jobs).  Chunk c = columns [c * chunk_cols, ...) belongs to GPU c % world."""
    lib = load_library()
    part = Partition(i0=0, j0=0, i1=m, j1=n)
    part.reserved[1] = chunk_cols
    info = ChainInfo()
    if lib.b200_chain_plan(C.byref(part), world, C.byref(info)) != 0:
        raise B200Error("b200_chain_plan: bad arguments")
    return {k: getattr(info, k) for k in ("chunks", "chunk_cols", "chunks_per_gpu", "max_strips", "max_jobs")}


def chain_chunks(n, world, chunk_cols=0):
    """[(j0, j1, owner)] of every chunk of an n-column chained partition."""
    if chunk_cols < 0:
        b = sorted(set(n * r // world for r in range(world + 1)))
    else:
        w = chain_plan(1, n, world, chunk_cols)["chunk_cols"]
        b = list(range(0, n, w)) + [n]
    return [(b[c], b[c + 1], c % world) for c in range(len(b) - 1)]


def merge_best(bests):
    """Global best of per-slice bests (score, i, j): highest score, then smallest i, then smallest j
    (C/common/BestScoreList.hpp:30-38)."""
    cand = [tuple(b) for b in bests if b is not None and b[1] >= 0]
    if not cand:
        return (-INF, -1, -1)
    return max(cand, key=lambda s: (s[0], -s[1], -s[2]))


class B200Error(RuntimeError):
    pass


def _as_u8(seq):
    if isinstance(seq, (bytes, bytearray)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    return np.ascontiguousarray(seq, dtype=np.uint8)


class Aligner:
    """Thin OO wrapper over one b200_handle."""

    def __init__(self, device=0, kernel=KERNEL_AUTO, warps_per_sm=0):
        self.lib = load_library()
        cfg = Config(device=device, kernel=kernel, warps_per_sm=warps_per_sm)
        self.h = C.c_void_p()
        rc = self.lib.b200_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise B200Error("b200_create failed: " + self.lib.b200_last_error(None).decode())
        self._seqs = None

    def close(self):
        if self.h:
            self.lib.b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise B200Error(f"{what} failed: {self.lib.b200_last_error(self.h).decode()}")

    def set_sequences(self, s0, s1):
        a, b = _as_u8(s0), _as_u8(s1)
        self._seqs = (a, b)
        self._check(self.lib.b200_set_sequences(self.h, a.ctypes.data, a.size, b.ctypes.data, b.size), "b200_set_sequences")

    def align_partition(self, i0=0, j0=0, i1=None, j1=None, recurrence=SMITH_WATERMAN, first_row_init=INIT_ZEROES,
                        first_col_init=INIT_ZEROES, first_row=None, first_col=None, special_row_interval=0,
                        block_height=0, want_special_rows=False, want_last_row=False, want_last_column=False,
                        want_best_score=True, prune=False, use_callbacks=True, mgpu=False, super_i1=None, super_j1=None,
                        chunk_cols=0, group=None):
        """Run b200_align_partition.  first_row / first_col: CELL arrays INCLUDING the corner as element 0
        (n+1 / m+1 cells), used when the init type is INIT_CUSTOM (or to feed gaps through the callback path)."""
        a, b = self._seqs
        i1 = a.size if i1 is None else i1
        j1 = b.size if j1 is None else j1
        part = Partition(i0=i0, j0=j0, i1=i1, j1=j1, recurrence=recurrence, first_row_init=first_row_init,
                         first_col_init=first_col_init, special_row_interval=special_row_interval,
                         block_height=block_height, want_special_rows=int(want_special_rows),
                         want_last_row=int(want_last_row), want_last_column=int(want_last_column),
                         want_best_score=int(want_best_score), prune=int(prune), super_i1=i1 if super_i1 is None else super_i1,
                         super_j1=j1 if super_j1 is None else super_j1)
        if mgpu:
            part.reserved[0] = 1          # B200_MGPU_CHAIN: the partition is the WHOLE one, this rank aligns its chunks
            part.reserved[1] = chunk_cols
        out = {"rows": {}, "row_first": {}, "last_column": [], "scores": []}
        pos = {"row": 0, "col": 0}

        def recv(kind, src):
            def fn(_ctx, buf, n):
                dst = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_int)), shape=(n * 2,)).view(CELL)
                if src is not None:
                    dst[:] = src[pos[kind]:pos[kind] + n]
                else:
                    p0 = pos[kind]
                    k = np.arange(p0, p0 + n, dtype=np.int64)
                    t = first_row_init if kind == "row" else first_col_init
                    if t == INIT_ZEROES:
                        dst["h"] = 0
                    else:
                        dst["h"] = np.where(k == 0, 0, -2 * k - (3 if t == INIT_GAPS else 0))
                    dst["x"] = -INF
                pos[kind] += n
            return RECV_FN(fn)

        def disp_row(_ctx, i, buf, n):
            arr = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_int)), shape=(n * 2,)).view(CELL).copy()
            out["rows"].setdefault(i, []).append(arr)

        def disp_col(_ctx, j, buf, n):
            arr = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_int)), shape=(n * 2,)).view(CELL).copy()
            out["last_column"].append(arr)

        def disp_score(_ctx, s):
            out["scores"].append((s.score, s.i, s.j))

        cbs = None
        keep = []
        if use_callbacks:
            keep = [recv("row", first_row), recv("col", first_col), DISP_FN(disp_row), DISP_FN(disp_col),
                    SCORE_FN(disp_score), CONT_FN(lambda _c: 1)]
            cbs = Callbacks(None, *keep)
        res = Result()
        if group is not None:
            part.reserved[1] = chunk_cols
            rc = self.lib.b200_group_align_partition(group, C.byref(part), C.byref(cbs) if cbs is not None else None, C.byref(res))
            if rc != 0:
                raise B200Error("b200_group_align_partition failed: " + self.lib.b200_group_last_error(group).decode())
        else:
            rc = self.lib.b200_align_partition(self.h, C.byref(part), C.byref(cbs) if cbs is not None else None, C.byref(res))
            self._check(rc, "b200_align_partition")
        out["best"] = (res.best.score, res.best.i, res.best.j)
        out["cells"] = res.cells
        out["cells_total"] = res.cells_total
        out["device_ms"] = res.device_ms
        out["strips"] = res.strips
        out["kernel_launches"] = res.kernel_launches
        out["kernel_used"] = res.kernel_used
        out["chunks"] = res.reserved[0]
        out["chunk_cols"] = res.reserved[1]
        out["warp_busy"] = res.reserved[2] / 1000.0
        out["warps"] = res.reserved[3]
        out["rows"] = {i: np.concatenate(v) for i, v in out["rows"].items()}
        out["last_column"] = np.concatenate(out["last_column"]) if out["last_column"] else np.zeros(0, CELL)
        return out

    # ---- multi-GPU chain -----------------------------------------------------------------------------
    def mgpu_setup(self, dist, rank, world, max_rows, max_cols, chunk_cols=0):
        """Export this rank's exchange block (sized for chained partitions of up to max_rows x max_cols with the given
        chunk width), all-gather the IPC handles over torch.distributed (dist=None: world 1, the GPU is its own
        neighbour), map the peers."""
        plan = chain_plan(max_rows, max_cols, world, chunk_cols)
        mine = (C.c_ubyte * 64)()
        self._check(self.lib.b200_mgpu_export(self.h, max_rows, plan["max_jobs"], mine), "b200_mgpu_export")
        handles = [bytes(mine)]
        if dist is not None:
            handles = [None] * world
            dist.all_gather_object(handles, bytes(mine))
        blob = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
        self._check(self.lib.b200_mgpu_connect(self.h, rank, world, blob), "b200_mgpu_connect")

    def last_chain_result(self):
        r = Result()
        self._check(self.lib.b200_last_chain_result(self.h, C.byref(r)), "b200_last_chain_result")
        return dict(best=(r.best.score, r.best.i, r.best.j), cells=r.cells, cells_total=r.cells_total, device_ms=r.device_ms,
                    warp_busy=r.reserved[2] / 1000.0, warps=r.reserved[3])

    # ---- diag primitives -------------------------------------------------------------------------------
    def diag_begin(self, part: Partition, split, block_height):
        sp = np.ascontiguousarray(split, dtype=np.int32)
        self._check(self.lib.b200_diag_begin(self.h, C.byref(part), sp.size - 1, sp.ctypes.data, block_height), "b200_diag_begin")

    def diag_set_first_row(self, cells, j):
        c = np.ascontiguousarray(cells, dtype=CELL)
        self._check(self.lib.b200_diag_set_first_row(self.h, c.ctypes.data, j, c.size), "b200_diag_set_first_row")

    def diag_set_first_column(self, cells, i, n):
        c = np.ascontiguousarray(cells, dtype=CELL)
        self._check(self.lib.b200_diag_set_first_column(self.h, c.ctypes.data, i, n), "b200_diag_set_first_column")

    def diag_process(self, diagonal, wl, wr):
        self._check(self.lib.b200_diag_process(self.h, diagonal, wl, wr), "b200_diag_process")

    def diag_get_row(self, j, n):
        out = np.zeros(n, CELL)
        self._check(self.lib.b200_diag_get_row(self.h, j, n, out.ctypes.data), "b200_diag_get_row")
        return out

    def diag_get_last_column(self, i, n):
        out = np.zeros(n, CELL)
        self._check(self.lib.b200_diag_get_last_column(self.h, i, n, out.ctypes.data), "b200_diag_get_last_column")
        return out

    def diag_get_block_scores(self, B):
        out = np.zeros(B, np.dtype([("i", "<i4"), ("j", "<i4"), ("score", "<i4")]))
        self._check(self.lib.b200_diag_get_block_scores(self.h, out.ctypes.data), "b200_diag_get_block_scores")
        return out

    def diag_clear_pruned(self, j0, j1):
        self._check(self.lib.b200_diag_clear_pruned(self.h, j0, j1), "b200_diag_clear_pruned")

    def diag_end(self):
        self._check(self.lib.b200_diag_end(self.h), "b200_diag_end")

    def match_last_column(self, buffer, base, goal):
        bu = np.ascontiguousarray(buffer, dtype=CELL)
        ba = np.ascontiguousarray(base, dtype=CELL)
        m = Match()
        self._check(self.lib.b200_match_last_column(self.h, bu.ctypes.data, ba.ctypes.data, bu.size, goal, C.byref(m)), "b200_match_last_column")
        return dict(found=bool(m.found), k=m.k, score=m.score, type=m.type)

    # ---- stage 4 ---------------------------------------------------------------------------------------
    def stage4_round(self, points, max_partition=16):
        """One batched split round: returns XPOINT[n]; entry k is the midpoint of partition (k-1, k), type -1 = none."""
        pts = np.ascontiguousarray(points, dtype=XPOINT)
        out = np.zeros(pts.size, XPOINT)
        self._check(self.lib.b200_stage4_round(self.h, pts.ctypes.data, pts.size, max_partition, out.ctypes.data), "b200_stage4_round")
        return out

    def stage4(self, points, max_partition=16):
        """All rounds + merges until the largest partition is <= max_partition (the reference's stage 4)."""
        pts = np.ascontiguousarray(points, dtype=XPOINT)
        cap = max(4 * (int(pts["i"].max() - pts["i"].min()) + int(pts["j"].max() - pts["j"].min())) + 64, 4 * pts.size)
        out = np.zeros(cap, XPOINT)
        n_out = C.c_int()
        self._check(self.lib.b200_stage4(self.h, pts.ctypes.data, pts.size, max_partition, out.ctypes.data, cap, C.byref(n_out)), "b200_stage4")
        return out[:n_out.value].copy()

    # ---- stage 5 ---------------------------------------------------------------------------------------
    def stage5(self, points):
        """Batched traceback of the partitions between consecutive crosspoints.  Returns (ops, op_off, op_len, stats):
        partition k owns ops[op_off[k] : op_off[k] + op_len[k]], one byte per step from its bottom-right corner."""
        pts = np.ascontiguousarray(points, dtype=XPOINT)
        cap = int(pts["i"][-1] - pts["i"][0]) + int(pts["j"][-1] - pts["j"][0])
        ops = np.full(cap + 1, 255, np.uint8)
        op_len = np.zeros(pts.size, np.int32)
        st = S5Stats()
        self._check(self.lib.b200_stage5(self.h, pts.ctypes.data, pts.size, ops.ctypes.data, cap, op_len.ctypes.data, C.byref(st)), "b200_stage5")
        off = np.zeros(pts.size, np.int64)
        off[1:] = (pts["i"][:-1].astype(np.int64) - int(pts["i"][0])) + (pts["j"][:-1].astype(np.int64) - int(pts["j"][0]))
        return ops[:cap], off, op_len, {k: getattr(st, k) for k, _ in S5Stats._fields_}

    def processed_cells(self):
        return self.lib.b200_processed_cells(self.h)

    def kernel_launches(self):
        return self.lib.b200_kernel_launches(self.h)


class Group:
    """Several GPUs driven by one process (b200_group_*): the multi-GPU mode of build/cudalign.  align_partition()
    has the semantics of the single-GPU call (whole rows, whole last column, merged best)."""

    def __init__(self, devices, max_rows, max_cols, chunk_cols=0, kernel=KERNEL_AUTO):
        self.lib = load_library()
        self.devices = list(devices)
        plan = chain_plan(max_rows, max_cols, len(self.devices), chunk_cols)
        dev = (C.c_int * len(self.devices))(*self.devices)
        cfg = Config(device=0, kernel=kernel, warps_per_sm=0)
        self.g = C.c_void_p()
        rc = self.lib.b200_group_create(dev, len(self.devices), C.byref(cfg), max_rows, plan["max_jobs"], C.byref(self.g))
        if rc != 0:
            raise B200Error("b200_group_create failed: " + self.lib.b200_group_last_error(None).decode())
        # a non-owning Aligner view of rank 0 reuses the callback plumbing of Aligner.align_partition
        self._view = Aligner.__new__(Aligner)
        self._view.lib = self.lib
        self._view.h = C.c_void_p()            # not owned: never destroyed through the view
        self._view._seqs = None

    def close(self):
        if self.g:
            self.lib.b200_group_destroy(self.g)
            self.g = C.c_void_p()

    def set_sequences(self, s0, s1):
        a, b = _as_u8(s0), _as_u8(s1)
        self._view._seqs = (a, b)
        rc = self.lib.b200_group_set_sequences(self.g, a.ctypes.data, a.size, b.ctypes.data, b.size)
        if rc != 0:
            raise B200Error("b200_group_set_sequences failed: " + self.lib.b200_group_last_error(self.g).decode())

    def align_partition(self, **kw):
        return self._view.align_partition(group=self.g, **kw)

    def rank_results(self):
        out = []
        for r in range(len(self.devices)):
            res = Result()
            self.lib.b200_group_rank_result(self.g, r, C.byref(res))
            out.append(dict(best=(res.best.score, res.best.i, res.best.j), cells=res.cells, cells_total=res.cells_total, device_ms=res.device_ms,
                            warp_busy=res.reserved[2] / 1000.0, warps=res.reserved[3]))
        return out
