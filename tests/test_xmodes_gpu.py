"""Drop-in parity in the CLI modes the other pipeline tests do not visit: all 25 --alignment-edges combinations (local, global and
the semi-global kinds, C/libmasa/libmasa.cpp:258-264) and --gpus with a chunk width far below the automatic one.

Both groups come from the differential fuzzing on the SIMT emulation (tests/emu/fuzz_pipeline.py, profiles/r02_emu_fuzz.txt), which
found two refusals of build/cudalign --gpus -- empty sequences handed over by MASA-Core for the degenerate stages of a semi-global
alignment ("b200_chain_plan failed"), and exchange blocks planned for the automatic chunk width while the B200_CHAIN_CHUNK
override was in force -- and --dump-blocks failing on the fast path.  The file is named to be collected last: it was written after the round's last GPU run and has run on the
emulation only (tests/test_emu_cpu.py), so under `pytest -x` it cannot hide the tests that have run on the B200."""
import os
import sys

import pytest

import oracle_lib as O  # noqa: F401
import test_pipeline_gpu as P

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import synth  # noqa: E402

pytestmark = pytest.mark.gpu

EDGES = [s + e for s in "*123+" for e in "*123+"]
CHAIN_ENV = {"B200_GROUP_WARPS_PER_SM": "8", "B200_GROUP_MIN_CELLS": "0", "B200_CHAIN_CHUNK": "700", "B200_WATCHDOG_S": "30"}


@pytest.fixture(scope="module")
def pair(tmp_path_factory):
    d = tmp_path_factory.mktemp("xmodes")
    a, b = synth.make_pair(5200, 4700, [(600, 4500)], 0.05, 0.02, 0.02, 1, 17)
    fa, fb = str(d / "A.fa"), str(d / "B.fa")
    synth.write_fasta(fa, a, "A")
    synth.write_fasta(fb, b, "B")
    return fa, fb


@pytest.mark.parametrize("edges", EDGES)
def test_alignment_edges_match_reference(tmp_path, tmp_path_factory, pair, edges):
    """Every start/end rule through the fast path and through stage 1 on a two-rank chain: crosspoint files, alignment.00.bin /
    .txt and special rows identical to the reference's CPU run."""
    P._need_binaries()
    fa, fb = pair
    extra = [f"--alignment-edges={edges}", "--no-block-pruning", "--disk-size=1M"]
    w_ref = P._ref_run(tmp_path_factory, "edges_" + edges.replace("*", "s").replace("+", "p"), fa, fb, extra)
    w_new = str(tmp_path / "new")
    P._run(P.CUDALIGN, fa, fb, w_new, extra)
    P._compare(w_ref, w_new, True)
    w_chain = str(tmp_path / "chain")
    P._run(P.CUDALIGN, fa, fb, w_chain, extra + ["--gpus=0,0"], env=CHAIN_ENV)
    P._compare(w_ref, w_chain, True)


def test_multi_gpu_pipeline_with_narrow_chunks(tmp_path, tmp_path_factory):
    """--gpus with a chunk width far below the automatic one: many more chain jobs per GPU than the default plan has.  The
    adapter must size the exchange blocks for the width it is going to use."""
    P._need_binaries()
    m, n = 8192, 20000
    a, b = synth.make_pair(m, n, [(1000, 7000)], 0.05, 0.02, 0.02, 0, 13)
    fa, fb = str(tmp_path / "A.fa"), str(tmp_path / "B.fa")
    synth.write_fasta(fa, a, "A")
    synth.write_fasta(fb, b, "B")
    extra = ["--no-block-pruning", "--disk-size=4M"]
    w_new = str(tmp_path / "new")
    w_ref = P._ref_run(tmp_path_factory, "narrow_chunks", fa, fb, extra)
    env = dict(CHAIN_ENV)
    env["B200_CHAIN_CHUNK"] = "32"
    P._run(P.CUDALIGN, fa, fb, w_new, extra + ["--gpus=0,0"], env=env)
    P._compare(w_ref, w_new, True)
    stats = open(os.path.join(w_new, "statistics.ALIGNER")).read()
    assert "on the multi-GPU chain" in stats and " 0 of them on the multi-GPU chain" not in stats, "stage 1 did not take the chain"


@pytest.mark.parametrize("mode", ["single", "gpus"])
def test_dump_blocks_takes_the_per_diagonal_path(tmp_path, tmp_path_factory, pair, mode):
    """--dump-blocks wants the score of every block of the reference's grid (AlignerManager.cpp:418-422, sw_stage1.cpp:310-314): the
    whole-partition kernel has no such artefact, so the adapter must fall back to the per-diagonal path -- pruning_dump.txt and the
    alignment (which embeds it, sw_stage5.cpp:456) identical to the reference's.  (Found on the emulation: the fast path exited 1.)"""
    P._need_binaries()
    fa, fb = pair
    extra = ["--dump-blocks", "--disk-size=1M"]
    w_ref = P._ref_run(tmp_path_factory, "dump_blocks", fa, fb, extra)
    w_new = str(tmp_path / "new")
    P._run(P.CUDALIGN, fa, fb, w_new, extra + (["--gpus=0,0"] if mode == "gpus" else []), env=CHAIN_ENV)
    P._compare(w_ref, w_new, False)
    assert open(os.path.join(w_ref, "pruning_dump.txt"), "rb").read() == open(os.path.join(w_new, "pruning_dump.txt"), "rb").read()
