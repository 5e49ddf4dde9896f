import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg():
    """The package directory is named uvic2.9_b200 (not an importable identifier); it is
    loaded under the module name uvic29_b200."""
    if "uvic29_b200" in sys.modules:
        return sys.modules["uvic29_b200"]
    path = os.path.join(ROOT, "uvic2.9_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(
        "uvic29_b200", path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["uvic29_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()
