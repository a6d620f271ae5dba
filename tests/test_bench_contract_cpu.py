"""CPU test (no GPU): bench.py keeps the driver's contract on the legs that run without a device.

* `--impl reference` prints exactly ONE line on stdout, a JSON object with the reference-arm keys;
* the product arm refuses to run without a CUDA device (no CPU fallback) and prints nothing on stdout."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "oracle_cpu_block")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (reference mount absent)")
def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "0.001", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GCUPS" and d["higher_is_better"] is True
    assert d["metric"] == "GCUPS (stage-1 SW, device-timed)"
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr


def test_context_legs_parse_and_degrade(tmp_path, monkeypatch):
    """The N = 1 context legs of the product arm never cost the bench line: the per-stage split is read from MASA-Core's own
    statistics file, and the reference-GPU-kernel leg reports `unavailable` where the two binaries do not exist.  With the
    CPU binaries of oracle/_ref standing in for the two executables the leg's whole flow (FASTA, both runs, ALIGN timer,
    crosspoint comparison) runs without a GPU."""
    import shutil
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench
    import synth
    st = tmp_path / "statistics"
    st.write_text("#  GLOBAL STATISTICS\n      SEQUENCES:       0.0560 (   1)  avg.:   0.0560\n         STAGE1:      29.1700 (   1)  avg.:  29.1700\n"
                  "         STAGE5:       0.3220 (   1)  avg.:   0.3220\n          TOTAL:      53.8340\n        Total: 53.8340\n")
    assert bench.stage_split(str(st)) == {"sequences": 0.056, "stage1": 29.17, "stage5": 0.322, "total": 53.834}
    assert bench.stage_split(str(tmp_path / "missing")) == {}
    monkeypatch.setattr(bench, "ROOT", str(tmp_path / "nowhere"))
    assert "unavailable" in bench.reference_gpu_leg(str(tmp_path))
    cpu_a, cpu_b = os.path.join(ROOT, "oracle", "_ref", "oracle_cpu"), os.path.join(ROOT, "oracle", "_ref", "oracle_cpu_block")
    if not (os.path.exists(cpu_a) and os.path.exists(cpu_b)):
        return
    fake = tmp_path / "root"
    (fake / "oracle" / "_ref").mkdir(parents=True)
    (fake / "build").mkdir()
    shutil.copy(cpu_a, fake / "oracle" / "_ref" / "cudalign_ref_gpu")
    shutil.copy(cpu_b, fake / "build" / "cudalign")
    monkeypatch.setattr(bench, "ROOT", str(fake))
    real = synth.make_config
    monkeypatch.setattr(synth, "make_config", lambda name, scale=1.0: real(name, 0.002 if name == "cfg1" else 0.0001))
    work = tmp_path / "work"
    work.mkdir()
    out = bench.reference_gpu_leg(str(work))
    assert len(out["pairs"]) == 2
    for pair in out["pairs"]:
        assert pair["same_result"] is True and pair["speedup"] > 0
        assert pair["reference_gpu"]["stage1_crosspoint"].count(",") == 3
