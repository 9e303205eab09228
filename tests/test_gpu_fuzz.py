"""Randomised shape sweep of EncoderLayer -> DecoderLayer (self-, cross-attention, FFN; d_k in {32, 64, 128}; lengths 1..520;
ragged batches) against the float64 oracle — tests/fuzz_layers.py run as a test."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [1, 2])
def test_random_layer_shapes_match_oracle(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_layers.py"), "8", str(seed)], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "fuzz ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
