import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with `-m gpu` on the GPU box")


@pytest.fixture(scope="session")
def lib_path():
    """Path of the in-tree shared library, built on demand (nvcc cross-compiles without a GPU)."""
    import speech_tranformer_pytorch_b200 as stb
    return stb.build()
