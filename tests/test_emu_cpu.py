"""The GPU parity tests, run where no GPU exists: on the SIMT emulation build of the library (tests/emu/).

tests/emu/ compiles the library's own sources -- masa-cudalign_b200/csrc/engine.cu with every kernel, the strip-chain protocol,
the multi-GPU chain, stage 4 and stage 5 -- for the host CPU: one fiber per CUDA thread, one OS thread per co-resident CTA, a
worker thread per stream, several emulated devices (tests/emu/cuda_runtime.h).  The result has the C ABI of libb200align.so,
so the `-m gpu` test files run on it UNCHANGED (in a subprocess, with the package's B200_LIB development switch pointing at
the emulation and, for build/cudalign, LD_PRELOAD): same inputs, same oracle, same bit-exact assertions.

This checks the C++ semantics of the device code and the protocols between warps, CTAs, streams and devices.  It is not a
CPU path of the product (nothing outside tests/ can load it, libb200align.so itself still refuses to run without a GPU:
tests/test_cabi_cpu.py) and not a substitute for the run on the B200: memory-ordering strength, convergence and timing are
properties of the hardware.  Cases are the small parametrisations of each GPU test file, sized for a few minutes in total."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import build_emu  # noqa: E402

pytestmark = pytest.mark.skipif(not build_emu.available(), reason="the SIMT emulation needs g++ on x86-64")


@pytest.fixture(scope="module")
def emu_lib():
    return build_emu.build()


F_PIPE = "tests/test_pipeline_gpu.py::"
GROUPS = {
    # name: (test files or node ids, -k expression, extra environment, LD_PRELOAD the emulation)
    "stage1": (["tests/test_stage1_gpu.py"],
               "test_sw_best_and_borders or test_special_rows or test_nw_global or test_custom_borders_subpartition "
               "or test_pruning_keeps_best_exact[20000-20000 or test_mixed_alphabet or test_chain_protocol_variants_are_exact[0] "
               "or test_chain_protocol_variants_are_exact[127] or test_packed_and_byte_sequences_agree or test_nw_border_with_minus_inf", {}, False),
    "diag": (["tests/test_diag_gpu.py", "tests/test_match_gpu.py", "tests/test_stage4_gpu.py", "tests/test_stage5_gpu.py"], None, {}, False),
    "chain": (["tests/test_chain_gpu.py", "tests/test_watchdog_gpu.py"],
              "test_self_chain_sw_matches_oracle[2049-1500-257 or test_self_chain_sw_matches_oracle[700-900 or test_self_chain_sw_matches_oracle[1-1 "
              "or test_self_chain_sw_matches_oracle[3000-3000 or test_self_chain_rearm or test_self_chain_nw_global "
              "or test_self_chain_custom_borders or test_group_on_one_device_sw[s16x2-4-700] or test_stuck_dependency", {}, False),
    "shuffle": (["tests/test_stage1_gpu.py", "tests/test_chain_gpu.py"],
                "test_sw_best_and_borders[2049-1500 or test_special_rows[9000-3000-s16x2] or test_self_chain_sw_matches_oracle[2049-1500-257 "
                "or test_self_chain_sw_matches_oracle[3000-3000 or test_group_on_one_device_sw[s16x2-4-700]", {"B200_EMU_SHUFFLE": "11"}, False),
    "pipeline": ([F_PIPE + "test_full_pipeline_matches_reference[fast-sw_3k]", F_PIPE + "test_full_pipeline_matches_reference[diag-sw_3k]",
                  F_PIPE + "test_full_pipeline_matches_reference[fast-sw_40k_pruning_ram]",
                  F_PIPE + "test_multi_gpu_pipeline_matches_reference[0,0-nw_global_20k]",
                  "tests/test_xmodes_gpu.py::test_multi_gpu_pipeline_with_narrow_chunks",
                  "tests/test_xmodes_gpu.py::test_alignment_edges_match_reference[*+]", "tests/test_xmodes_gpu.py::test_alignment_edges_match_reference[12]",
                  "tests/test_xmodes_gpu.py::test_alignment_edges_match_reference[3+]", "tests/test_xmodes_gpu.py::test_alignment_edges_match_reference[21]",
                  "tests/test_xmodes_gpu.py::test_dump_blocks_takes_the_per_diagonal_path"],
                 None, {}, True),
}


def _have_pipeline_binaries():
    return os.path.exists(os.path.join(ROOT, "build", "cudalign")) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "oracle_cpu"))


def _launch(emu_lib, cmd, extra_env=None, preload=False, timeout=1500):
    env = dict(os.environ)
    env["B200_LIB"] = emu_lib
    env["B200_TEST_EMULATION"] = "1"
    if preload:
        env["LD_PRELOAD"] = emu_lib          # build/cudalign links libb200align.so: the emulation's b200_* symbols take precedence
    env.update(extra_env or {})
    return subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)


@pytest.fixture(scope="module")
def runs(emu_lib):
    """Every group of this module is a subprocess (the package binds one library per process); they are started together, three
    at a time, and each test below waits for its own."""
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(3)
    out = {}
    for name, (files, k, extra, preload) in GROUPS.items():
        if name == "pipeline" and not _have_pipeline_binaries():
            continue
        cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", "--timeout", "900", *files] + (["-k", k] if k else [])
        out[name] = pool.submit(_launch, emu_lib, cmd, extra, preload)
    out["torchrun"] = pool.submit(
        _launch, emu_lib,
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", str(29700 + os.getpid() % 200), os.path.join(ROOT, "tests", "mgpu_check.py")],
        dict(B200_EMU_SHM="1", B200_EMU_DEVICES="2", MGPU_CHECK_BACKEND="gloo", B200_WATCHDOG_S="120", MGPU_CASES="sw_small_s16,sw_small_s32,nw_global_s32"))
    out["fuzz"] = pool.submit(_launch, emu_lib, [sys.executable, os.path.join(ROOT, "tests", "emu", "fuzz_emu.py"), "--cases", "150", "--seed", "21",
                                                 "--max-side", "2500"])
    yield out
    pool.shutdown(wait=True)


def _passed(fut):
    """Number of GPU-marked tests that passed in a finished pytest subprocess; none may fail or skip itself."""
    r = fut.result()
    tail = r.stdout[-3000:]
    last = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
    assert r.returncode == 0 and "failed" not in last, tail
    assert not re.search(r"\d+ skipped", last), "a selected test skipped itself:\n" + tail
    m = re.search(r"(\d+) passed", last)
    return int(m.group(1)) if m else 0


def test_emulation_is_what_runs(emu_lib):
    """The subprocesses below really load the emulation: it exports the emulation-only probe, the product library does not."""
    import ctypes
    lib = ctypes.CDLL(emu_lib)
    assert lib.b200_emu_is_emulation() == 1
    prod = os.path.join(ROOT, "masa-cudalign_b200", "libb200align.so")
    if os.path.exists(prod):
        assert not hasattr(ctypes.CDLL(prod), "b200_emu_is_emulation")


def test_package_refuses_the_emulation_outside_a_test_run(emu_lib):
    """Pointing the package at the emulation is not a way to run the product on a CPU: without the test switch it refuses."""
    env = {k: v for k, v in os.environ.items() if k != "B200_TEST_EMULATION"}
    env["B200_LIB"] = emu_lib
    code = "from __graft_entry__ import load_package; load_package().load_library()"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU path" in r.stdout, r.stdout[-2000:]


def test_stage1_whole_partition_path(runs):
    """Both strip kernels against the oracle: best cell, last row / column, special rows, NW, custom borders, N / IUPAC
    bytes, packed sequences, -INF borders, on-device pruning, one non-default variant of the strip-chain protocol."""
    assert _passed(runs["stage1"]) >= 30


def test_diag_primitives_matcher_stage4_stage5(runs):
    """Every diag primitive (the reference's per-diagonal contract), the device goal matcher, the batched stage-4 split
    against the reference's crosspoint_04 and the stage-5 traceback against the reference's alignments."""
    assert _passed(runs["diag"]) >= 60


def test_chain_on_one_and_several_emulated_devices(runs):
    """The block-cyclic chain: a device that is its own neighbour, 4 ranks of a group, NW, custom borders, re-arming with other
    sequences; the watchdog reports a stuck dependency."""
    assert _passed(runs["chain"]) >= 16


def test_scheduling_order_does_not_matter(runs):
    """Stage-1 and chain cases with the fibers of every CTA visited in a random order that changes each pass: the results may not
    depend on which warp or lane runs first (a cheap search for protocol races)."""
    assert _passed(runs["shuffle"]) >= 7


def test_drop_in_binary_full_pipeline(runs):
    """build/cudalign (B200Aligner + the reference's unmodified MASA-Core, stages 1-6, stage 4 and 5 substitutes) against the
    reference's CPU run: crosspoint files, alignment.00.bin / .txt, special rows -- fast path, diag path, stage 1 on a chain of
    two ranks, semi-global --alignment-edges modes, a chunk width far below the automatic one."""
    if "pipeline" not in runs:
        pytest.skip("build/cudalign or oracle/_ref/oracle_cpu not built (they need the reference mount at build time)")
    assert _passed(runs["pipeline"]) == 11


def test_one_process_per_device_chain_under_torchrun(runs):
    """tests/mgpu_check.py, the script tests/test_mgpu_gpu.py launches on 2 / 4 / 8 real GPUs, with two PROCESSES of one emulated
    device each (gloo instead of NCCL; the exchange blocks are shared memory mapped through the emulation's CUDA IPC): border
    stores, event words, job queues and the running best cross the process border as they cross NVLink."""
    p = runs["torchrun"].result()
    assert p.returncode == 0, p.stdout[-3000:]
    assert p.stdout.count(": OK") == 3 and "MISMATCH" not in p.stdout, p.stdout[-3000:]


def test_fuzz_sweep_against_the_oracle(runs):
    """A short fixed-seed sweep of tests/emu/fuzz_emu.py: random shapes around the tile boundaries, both kernels, SW / NW, every
    border kind (also offset by millions), N / IUPAC bytes, pruning, self-chain and groups on distinct emulated devices -- bit-exact
    against the oracle.  The sweep also counts s16 overflows of the packed arithmetic: the only ones this seed produces are the
    16 of case 125 (dead lanes below the last row of a partial strip whose frame is ~ -32770: DESIGN.md section 7, item 7)."""
    r = runs["fuzz"].result()
    assert r.returncode == 0, r.stdout[-3000:]
    assert "fuzz: 150 cases, 0 mismatches, 16 s16 overflow events" in r.stdout, r.stdout[-1500:]
    assert r.stdout.count("s16 overflow events]") == 1 and "case 125: 17155x512 s16x2 NW" in r.stdout
