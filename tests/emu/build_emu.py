"""TEST INFRASTRUCTURE: builds tests/emu/_build/libb200align_emu.so -- the library's own sources (masa-cudalign_b200/csrc/engine.cu
and everything it includes) compiled by g++ for the host CPU against the SIMT emulation in this directory (cuda_runtime.h,
emu_ptx.h, emu_runtime.cpp).  Same C ABI as libb200align.so; loaded only by tests (tests/test_emu_cpu.py, through the
B200_LIB development switch of the package).  `python tests/emu/build_emu.py` builds it by hand."""
import os
import platform
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "masa-cudalign_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libb200align_emu.so")


def available():
    """The fiber switch is x86-64 assembly; everything else is portable C++17."""
    return platform.machine() in ("x86_64", "AMD64") and shutil.which(os.environ.get("CXX", "g++")) is not None


def sources():
    src = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    src += [os.path.join(HERE, f) for f in ("cuda_runtime.h", "emu_ptx.h", "emu_runtime.cpp")]
    src.append(os.path.join(ROOT, "include", "b200align.h"))
    return src


def build(force=False):
    if not force and os.path.exists(OUT):
        t = os.path.getmtime(OUT)
        if all(os.path.getmtime(s) <= t for s in sources()):
            return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cxx = os.environ.get("CXX", "g++")
    tmp = OUT + ".tmp%d" % os.getpid()
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-pthread", "-DB200_EMU", "-Wno-unknown-pragmas",
                           "-I", HERE, "-I", CSRC, "-x", "c++", os.path.join(CSRC, "engine.cu"), os.path.join(HERE, "emu_runtime.cpp"),
                           "-o", tmp], cwd=ROOT)
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
