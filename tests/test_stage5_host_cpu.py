"""CPU test (no GPU) of the stage-5 traceback the GPU threads execute: csrc/stage5.cuh's s5_partition is __host__
__device__, tests/harness/s5_host.cu builds it for the host with nvcc, and the walk is compared with
  * the reference's own stage 5 (golden alignment.00.bin contents, tests/golden/stage5_runs.json),
  * the oracle restatement on random partition chains (all nine start/end type pairs, ragged sizes, both storage variants).
The kernels themselves run in tests/test_stage5_gpu.py."""
import ctypes as C
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "stage5_runs.json")))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    nvcc = shutil.which(os.environ.get("NVCC", "nvcc"))
    if nvcc is None:
        pytest.skip("nvcc not available")
    so = str(tmp_path_factory.mktemp("s5") / "s5_host.so")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                           "-o", so, os.path.join(ROOT, "tests", "harness", "s5_host.cu")])
    lib = C.CDLL(so)
    lib.s5_host_walk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return lib


def host_walk(lib, a, b, pts, force_global):
    pts = np.ascontiguousarray(pts, dtype=O.XPOINT)
    cap = int(pts["i"][-1] - pts["i"][0]) + int(pts["j"][-1] - pts["j"][0])
    ops = np.full(cap + 1, 255, np.uint8)
    op_len = np.zeros(pts.size, np.int32)
    tot = np.zeros(5, np.int32)
    assert lib.s5_host_walk(a.ctypes.data, b.ctypes.data, pts.ctypes.data, pts.size, ops.ctypes.data, op_len.ctypes.data, tot.ctypes.data, int(force_global)) == 0
    return ops[:cap], op_len, tot


def check_against_oracle(lib, a, b, pts, force_global):
    ops_o, off, len_o, st = O.stage5(a, b, pts)
    ops_h, len_h, tot = host_walk(lib, a, b, pts, force_global)
    dev = len_h >= 0                      # partitions that reach the device
    dev[0] = False
    assert np.array_equal(len_h[dev], len_o[dev])
    for k in np.nonzero(dev)[0]:
        s = slice(int(off[k]), int(off[k]) + int(len_o[k]))
        assert np.array_equal(ops_h[s], ops_o[s]), f"partition {k}"
    # totals: the oracle's sum minus what the pure-gap partitions contribute (walked on the host in the product as well)
    sub = np.ascontiguousarray(pts, dtype=O.XPOINT)
    gap_tot = np.zeros(5, np.int64)
    for k in np.nonzero(~dev)[0]:
        if k == 0:
            continue
        _o, _f, _l, s1 = O.stage5(a, b, sub[k - 1:k + 1])
        gap_tot += [s1["score"], s1["matches"], s1["mismatches"], s1["gap_open"], s1["gap_ext"]]
    want = np.array([st["score"], st["matches"], st["mismatches"], st["gap_open"], st["gap_ext"]]) - gap_tot
    assert np.array_equal(tot, want)


@pytest.mark.parametrize("force_global", [False, True], ids=["local<=32", "global"])
@pytest.mark.parametrize("name", sorted(GOLD))
def test_device_function_reproduces_reference_stage5(host_lib, name, force_global):
    g = GOLD[name]
    ge = g["generator"]
    a, b = synth.make_pair(ge["m"], ge["n"], [tuple(s) for s in ge["segments"]], ge["p_s"], ge["p_d"], ge["p_i"], ge["K"], ge["seed"])
    pts = O.golden_points(g["crosspoint_04"])
    check_against_oracle(host_lib, a, b, pts, force_global)
    # and straight against the reference's stored gap lists: device steps for the real partitions, host rule for pure gaps
    ops_o, off, len_o, _st = O.stage5(a, b, pts)
    ops_h, len_h, _t = host_walk(host_lib, a, b, pts, force_global)
    ops = ops_o.copy()
    for k in range(1, pts.size):
        if len_h[k] >= 0:
            ops[int(off[k]):int(off[k]) + int(len_h[k])] = ops_h[int(off[k]):int(off[k]) + int(len_h[k])]
    g0, g1 = O.stage5_gaps(pts, ops, off, np.where(len_h >= 0, len_h, len_o))
    assert g0 == sorted(g["alignment"]["gaps0"]) and g1 == sorted(g["alignment"]["gaps1"])


@pytest.mark.parametrize("seed", range(6))
def test_device_function_on_random_partition_chains(host_lib, seed):
    """Arbitrary (not optimal-path) crosspoint chains: every start/end type pair, sizes 1..70, unrelated and similar
    sequences -- the walk must still be the reference's, step for step (ties, -INF borders, leftovers)."""
    rng = np.random.default_rng(1000 + seed)
    m = n = 6000
    a, b = synth.make_pair(m, n, [(0, m)], 0.10, 0.05, 0.05, 0, 500 + seed)
    if seed % 2:
        b = synth.ACGT[rng.integers(0, 2, size=n)]          # two-letter alphabet: many ties
        a = synth.ACGT[rng.integers(0, 2, size=m)]
    pts = [(0, 0, int(rng.integers(0, 3)), 0)]
    while True:
        i, j = pts[-1][0], pts[-1][1]
        di, dj = int(rng.integers(0, 71)), int(rng.integers(0, 71))
        if rng.random() < 0.6:
            di, dj = min(di, 16), min(dj, 16)
        if i + di > m or j + dj > n:
            break
        if di == 0 and dj == 0:
            continue
        pts.append((i + di, j + dj, int(rng.integers(0, 3)), 0))
    pts = np.array(pts, dtype=O.XPOINT)
    check_against_oracle(host_lib, a, b, pts, False)
    check_against_oracle(host_lib, a, b, pts, True)


def test_device_function_at_the_table_limit(host_lib):
    """Partitions at the reference's table limit (1024 on a side, H_MAX / W_MAX of sw_stage5.cpp:31-32), thin ones, and one cell."""
    rng = np.random.default_rng(9)
    m, n = 5000, 5200
    a, b = synth.make_pair(m, n, [(0, m)], 0.06, 0.03, 0.03, 0, 91)
    pts = [(0, 0, 0, 0)]
    for di, dj in [(1024, 1024), (1, 1), (1, 1024), (1024, 1), (33, 32), (32, 33), (700, 900), (1024, 1000)]:
        pts.append((pts[-1][0] + di, pts[-1][1] + dj, int(rng.integers(0, 3)), 0))
    pts = np.array(pts, dtype=O.XPOINT)
    check_against_oracle(host_lib, a, b, pts, False)
    check_against_oracle(host_lib, a, b, pts, True)
