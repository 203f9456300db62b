import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_package():
    """Import the product package.  Its directory name (nllssolver.jl_b200) contains a dot, so it is
    registered under the importable alias `nllssolver_jl_b200`."""
    name = "nllssolver_jl_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkgdir = os.path.join(ROOT, "nllssolver.jl_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkgdir, "__init__.py"), submodule_search_locations=[pkgdir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle
