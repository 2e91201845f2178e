import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `pytest -m gpu` under gpurun)")


@pytest.fixture(scope="session")
def sgpkg():
    return importlib.import_module("scrabble-gan_b200")


@pytest.fixture(scope="session")
def rt(sgpkg):
    """Session runtime on cuda:0 (GPU tests only)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    runtime = importlib.import_module("scrabble-gan_b200.runtime")
    r = runtime.Runtime(device=0, mode="fp32")
    runtime.set_runtime(r)
    return r
