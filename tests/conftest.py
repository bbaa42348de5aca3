import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg():
    """Import the product package. Its directory is named `wflow.jl_b200` (not a valid
    dotted module name), so it is loaded by path under the module name `wflow_jl_b200`."""
    name = "wflow_jl_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = os.path.join(ROOT, "wflow.jl_b200")
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()
