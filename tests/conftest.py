import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def load_pkg():
    """Import the hyphenated package directory fast-llama_b200/ as module `fast_llama_b200`."""
    if "fast_llama_b200" in sys.modules:
        return sys.modules["fast_llama_b200"]
    init = os.path.join(ROOT, "fast-llama_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("fast_llama_b200", init, submodule_search_locations=[os.path.dirname(init)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["fast_llama_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fl():
    return load_pkg()
