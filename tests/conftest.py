import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_package():
    """import ligero-prover_b200/ (hyphenated directory) as module `ligero_prover_b200`"""
    if "ligero_prover_b200" in sys.modules:
        return sys.modules["ligero_prover_b200"]
    path = os.path.join(ROOT, "ligero-prover_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("ligero_prover_b200", path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ligero_prover_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lgr():
    return load_package()


@pytest.fixture(scope="session")
def oracle():
    from oracle import lgo
    lgo.build()
    return lgo


_executors = {}


@pytest.fixture(scope="session")
def executor_factory(lgr):
    """cached Executor per k (contexts build twiddle tables; reuse them across tests)"""
    def make(k, l=None):
        key = (k, l)
        if key not in _executors:
            _executors[key] = lgr.make_executor(l if l is not None else max(k - 192, 1), k)
        ex = _executors[key]
        ex.use_torch_stream()
        return ex
    yield make
    for ex in _executors.values():
        ex.close()
    _executors.clear()
