import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _load_pkg():
    """The package directory is `masa-cudalign_b200/` (not an identifier): import it as masa_cudalign_b200."""
    name = "masa_cudalign_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "masa-cudalign_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    _load_pkg()


@pytest.fixture(scope="session")
def b200():
    return _load_pkg()


@pytest.fixture(scope="session")
def aligner(b200):
    import torch  # noqa: F401  (only to fail early with a clear message when no GPU is visible)
    a = b200.Aligner(device=0)
    yield a
    a.close()
