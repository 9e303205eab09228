"""Run every device self-test of libst_b200.so (tcgen05 building blocks) and print the relative errors."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401  (CUDA context)
import speech_tranformer_pytorch_b200 as stb  # noqa: E402

lib = stb._lib.load()
torch.zeros(1, device="cuda")
first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
bad = 0
for i in range(first, lib.st_selftest_count()):
    err = ctypes.c_double(-1)
    st = lib.st_selftest(i, ctypes.byref(err))
    ok = st == 0 and 0 <= err.value < 1e-4
    bad += not ok
    print(f"selftest {i:3d}: status {st} err {err.value:.3e} {'ok' if ok else 'FAIL ' + lib.st_last_error().decode()}", flush=True)
print("FAILED" if bad else "all ok", bad)
