"""End-to-end drop-in parity: build/cudalign (B200Aligner + unmodified MASA-Core stages 1-6) against the parity
oracle oracle/_ref/oracle_cpu (reference CPUBlockProcessor under the reference's AbstractDiagonalAligner policy)
on the same FASTA files.  Bit-exact artefacts: crosspoint files of stages 1-4, alignment.00.bin, alignment.00.txt
(minus the header lines that carry file names), and -- with pruning off -- the stage-1 special-row files."""
import filecmp
import os
import subprocess
import sys

import pytest

import oracle_lib as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import synth  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDALIGN = os.path.join(ROOT, "build", "cudalign")


def _need_binaries():
    if not (os.path.exists(CUDALIGN) and O.have_ref_binaries()):
        pytest.skip("build/cudalign or oracle/_ref binaries not built (need the reference mount at build time)")


def _run(exe, fa, fb, wd, extra, env=None):
    cmd = [exe, f"--work-dir={wd}", "--clear", "--verbose=0", *extra, fa, fb]
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1200, env=e)
    assert r.returncode == 0, f"{' '.join(cmd)}\nexit code {r.returncode}\n{r.stdout[-3000:]}"


_REF_CACHE = {}


def _ref_run(tmp_path_factory, name, fa, fb, extra):
    """The reference's CPU run of a case, once per session: the same pair + flags serve several parametrisations (fast /
    diag path, 2 / 4 ranks), and the reference needs about a minute of CPU for the 150k x 140k pairs."""
    key = (name, tuple(extra))
    if key not in _REF_CACHE:
        wd = str(tmp_path_factory.mktemp("ref_" + name) / "w")
        _run(os.path.join(O.REF_DIR, "oracle_cpu"), fa, fb, wd, extra)
        _REF_CACHE[key] = wd
    return _REF_CACHE[key]


def _files(d):
    out = {}
    for base, _dirs, names in os.walk(d):
        for n in names:
            p = os.path.join(base, n)
            out[os.path.relpath(p, d)] = p
    return out


def _body(path):
    return [l for l in open(path) if not (l.startswith("Query: ") or l.startswith("Sbjct: ")) or l[7:20].strip()[:1].isdigit()]


def _compare(w_ref, w_new, special_rows):
    xr, xn = _files(os.path.join(w_ref, "crosspoints")), _files(os.path.join(w_new, "crosspoints"))
    assert sorted(xr) == sorted(xn), (sorted(xr), sorted(xn))
    for k in sorted(xr):
        assert open(xr[k]).read() == open(xn[k]).read(), f"crosspoint file {k} differs"
    assert filecmp.cmp(os.path.join(w_ref, "alignment.00.bin"), os.path.join(w_new, "alignment.00.bin"), shallow=False), "alignment.00.bin differs"
    assert _body(os.path.join(w_ref, "alignment.00.txt")) == _body(os.path.join(w_new, "alignment.00.txt")), "alignment.00.txt differs"
    if special_rows:
        sr, sn = _files(os.path.join(w_ref, "special_rows", "stage.01.00")), _files(os.path.join(w_new, "special_rows", "stage.01.00"))
        assert sorted(sr) == sorted(sn), (sorted(sr)[:10], sorted(sn)[:10])
        for k in sorted(sr):
            assert filecmp.cmp(sr[k], sn[k], shallow=False), f"special row file {k} differs"
        return len(sr)
    return 0


CASES = [
    # name, m, n, homology, extra flags, compare special rows
    ("sw_3k", 3000, 2700, (500, 2500), [], False),
    ("sw_40k_nopruning_disk", 40000, 39979, (5000, 35000), ["--no-block-pruning", "--disk-size=4M"], True),
    ("sw_40k_pruning_ram", 40000, 39979, (5000, 35000), ["--ram-size=3M"], False),
    ("nw_global_20k", 20000, 21000, (0, 20000), ["--alignment-edges=++", "--disk-size=4M", "--no-block-pruning"], True),
    ("sw_150k_nopruning", 150000, 140000, (20000, 120000), ["--no-block-pruning", "--disk-size=40M"], True),
]


@pytest.mark.parametrize("name,m,n,hom,extra,sr", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("path", ["fast", "diag"])
def test_full_pipeline_matches_reference(tmp_path, tmp_path_factory, name, m, n, hom, extra, sr, path):
    _need_binaries()
    a, b = synth.make_pair(m, n, [hom], 0.05, 0.02, 0.02, 0, 11)
    fa, fb = str(tmp_path / "A.fa"), str(tmp_path / "B.fa")
    synth.write_fasta(fa, a, "A")
    synth.write_fasta(fb, b, "B")
    w_new = str(tmp_path / "new")
    w_ref = _ref_run(tmp_path_factory, name, fa, fb, extra)
    _run(CUDALIGN, fa, fb, w_new, extra + (["--no-fast-path"] if path == "diag" else []))
    nrows = _compare(w_ref, w_new, sr)
    if sr:
        assert nrows > 0


MGPU_CASES = [
    ("sw_40k_nopruning_disk", 40000, 39979, (5000, 35000), ["--no-block-pruning", "--disk-size=4M"], True),
    ("sw_40k_pruning_ram", 40000, 39979, (5000, 35000), ["--ram-size=3M"], False),
    ("nw_global_20k", 20000, 21000, (0, 20000), ["--alignment-edges=++", "--disk-size=4M", "--no-block-pruning"], True),
    ("sw_150k_pruning", 150000, 140000, (20000, 120000), ["--disk-size=40M"], False),
]


@pytest.mark.parametrize("name,m,n,hom,extra,sr", MGPU_CASES, ids=[c[0] for c in MGPU_CASES])
@pytest.mark.parametrize("gpus", ["0,0", "0,0,0,0"])
def test_multi_gpu_pipeline_matches_reference(tmp_path, tmp_path_factory, name, m, n, hom, extra, sr, gpus):
    """build/cudalign --gpus=...: stage 1 on the block-cyclic chain (here: several ranks sharing device 0; with real
    device lists the same code path runs over NVLink, tests/test_mgpu_gpu.py), stages 2-6 on the first GPU.  The
    artefacts must be the reference's, byte for byte -- special rows included: they are assembled from the chunks of all
    ranks before they reach MASA-Core's special-rows area."""
    _need_binaries()
    a, b = synth.make_pair(m, n, [hom], 0.05, 0.02, 0.02, 0, 11)
    fa, fb = str(tmp_path / "A.fa"), str(tmp_path / "B.fa")
    synth.write_fasta(fa, a, "A")
    synth.write_fasta(fb, b, "B")
    w_new = str(tmp_path / "new")
    w_ref = _ref_run(tmp_path_factory, name, fa, fb, extra)
    nranks = gpus.count(",") + 1
    _run(CUDALIGN, fa, fb, w_new, extra + [f"--gpus={gpus}"],
         env={"B200_GROUP_WARPS_PER_SM": str(16 // nranks), "B200_GROUP_MIN_CELLS": "0", "B200_CHAIN_CHUNK": "3000", "B200_WATCHDOG_S": "30"})
    sp = os.path.join(w_new, "statistics.ALIGNER")        # printFinalStatistics (libmasa.cpp:1336,1398)
    stats = open(sp).read() if os.path.exists(sp) else ""
    nrows = _compare(w_ref, w_new, sr)
    if sr:
        assert nrows > 0
    assert "on the multi-GPU chain" in stats and " 0 of them on the multi-GPU chain" not in stats, "stage 1 did not take the chain"

