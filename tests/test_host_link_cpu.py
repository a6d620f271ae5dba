"""CPU test (no GPU) of how build/cudalign is put together: the link-time substitutes (stage 4, stage 5, RAM special-row
allocator) are the definitions that ended up in the binary, they call the C ABI, and nothing of oracle/ (test
infrastructure) is linked or referenced.  Skipped where the binary cannot be built (no reference mount)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "cudalign")


def _nm():
    if not os.path.exists(EXE) or shutil.which("nm") is None:
        pytest.skip("build/cudalign not built here")
    return subprocess.check_output(["nm", "-C", EXE], text=True).splitlines()


def test_substitutes_are_linked_and_call_the_c_abi():
    sym = _nm()
    defined = {l.split(" ", 2)[2] for l in sym if len(l.split(" ", 2)) == 3 and l.split(" ", 2)[1] in "Tt"}
    undefined = {l.split()[-1] for l in sym if " U " in l}
    for name in ("stage4(Job*, int)", "stage5(Job*, int)", "SpecialRowRAM::initialize(bool, int)"):
        assert name in defined, name
    for name in ("b200_stage4", "b200_stage5", "b200_align_partition", "b200_group_align_partition", "b200_set_sequences"):
        assert name in undefined, f"{name} is not imported from libb200align.so"
    # exactly one definition of each substituted function made it into the binary (the archive's copy was not pulled on top)
    for name in ("stage5(Job*, int)", "SpecialRowRAM::initialize(bool, int)"):
        assert sum(1 for l in sym if l.endswith(" T " + name)) == 1


def test_product_binary_does_not_depend_on_oracle():
    _nm()
    needed = subprocess.check_output(["readelf", "-d", EXE], text=True) if shutil.which("readelf") else ""
    assert "oracle" not in needed
    assert "libb200align.so" in needed
    blob = open(EXE, "rb").read()
    assert b"oracle/_ref" not in blob and b"gotoh_oracle" not in blob
    lib = os.path.join(ROOT, "masa-cudalign_b200", "libb200align.so")
    assert b"gotoh_oracle" not in open(lib, "rb").read()
