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


@pytest.fixture(params=["fp16", "tf32"])
def engine(request):
    """Runs the test once per fp32 engine of the composite operators (functional.set_fp32_engine): "fp16" = fp32 boundary
    tensors with fp16 operands inside (ST_DTYPE_F32_H16, the default), "tf32" = TF32 operands throughout (ST_DTYPE_F32)."""
    from speech_tranformer_pytorch_b200 import functional as F
    prev = F.set_fp32_engine(request.param)
    yield request.param
    F.set_fp32_engine(prev)
