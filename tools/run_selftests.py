"""Run every libst_b200 device self-test in its own process (a trap in one must not poison the rest).
Usage: python tools/run_selftests.py [case ...]   -> prints one line per case, exit code 1 on any failure."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "speech-tranformer-pytorch_b200", "libst_b200.so")

def one(which: int) -> int:
    lib = ctypes.CDLL(LIB)
    lib.st_last_error.restype = ctypes.c_char_p
    if lib.st_device_check(0) != 0:
        print(f"case {which}: DEVICE {lib.st_last_error().decode()}"); return 2
    err = ctypes.c_double(-1)
    st = lib.st_selftest(which, ctypes.byref(err))
    ok = st == 0 and 0 <= err.value < 1e-4
    print(f"case {which}: status {st} rel_err {err.value:.3e} {'OK' if ok else 'FAIL ' + lib.st_last_error().decode()}", flush=True)
    return 0 if ok else 1

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        sys.exit(one(int(sys.argv[2])))
    lib = ctypes.CDLL(LIB)
    cases = [int(a) for a in sys.argv[1:]] or list(range(lib.st_selftest_count()))
    bad = 0
    for c in cases:
        try:
            r = subprocess.run([sys.executable, __file__, "--one", str(c)], capture_output=True, text=True, timeout=90)
            out = (r.stdout + r.stderr).strip()
            print(out if out else f"case {c}: no output rc={r.returncode}", flush=True)
            bad += r.returncode != 0
        except subprocess.TimeoutExpired:
            print(f"case {c}: TIMEOUT", flush=True); bad += 1
    print(f"selftests: {len(cases) - bad}/{len(cases)} passed")
    sys.exit(1 if bad else 0)
