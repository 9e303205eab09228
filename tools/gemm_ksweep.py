"""Per-tile cost model of the TF32 GEMM: sweep K at fixed M, N -> t = waves * (a + b*K).  The intercept a is the
per-tile epilogue / hand-off cost, the slope b the main-loop cost per k."""
import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
L = stb._lib; lib = L.load(); DEV = "cuda:0"
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)
def run(M, N, K, bn, cl, do_flush=True):
    lib.st_set_option(b"gemm_bn", bn); lib.st_set_option(b"gemm_cluster", cl)
    A = torch.randn(M, K, device=DEV); B = torch.randn(N, K, device=DEV); Cm = torch.zeros(M, N, device=DEV)
    ep = L.GemmEpilogue(bias=None, aux=None, ldaux=N, aux_mode=0, relu=0, round_tf32=1, k_splits=1, dropout_p=0.0, seed=7)
    ts = []
    for i in range(6):
        if do_flush: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.check(lib.st_gemm(0, A.data_ptr(), K, B.data_ptr(), K, Cm.data_ptr(), N, M, N, K, C.byref(ep), None)); e1.record()
        torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    t = min(ts) * 1e3
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn); waves = -(-tiles // 148)
    print(f"M{M} N{N} K{K:5d} bn{bn} cl{cl} flush{int(do_flush)}: {t:8.1f} us  {t / waves:6.2f} us/tile-wave  {2.0*M*N*K/t/1e6:6.1f} TFLOP/s", flush=True)
for bn, cl in [(256, 1), (256, 3), (128, 1)]:
    for K in [32, 64, 128, 256, 512, 1024, 2048]:
        run(32000, 2048, K, bn, cl)
run(32000, 2048, 512, 256, 1, do_flush=False); run(32000, 2048, 32, 256, 1, do_flush=False)
