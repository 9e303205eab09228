"""Time the tcgen05 TF32 GEMM on the headline shapes (CUDA events, L2 flushed between launches).
python tools/bench_gemm.py [--one M N K mode]"""
import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
L = stb._lib; lib = L.load(); DEV = "cuda:0"
def p(t): return None if t is None else t.data_ptr()
if os.environ.get("ST_GEMM_CLUSTER"): lib.st_set_option(b"gemm_cluster", int(os.environ["ST_GEMM_CLUSTER"]))
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)
def run(mode, M, N, K, splits=1, iters=5, bias=True, aux_mode=0, drop=0.0):
    A = torch.randn((K, M) if mode == 2 else (M, K), device=DEV); B = torch.randn((N, K) if mode == 0 else (K, N), device=DEV)
    Cm = torch.zeros(M, N, device=DEV); bvec = torch.randn(N, device=DEV) if bias and mode == 0 else None
    aux = torch.randn(M, N, device=DEV) if aux_mode else None
    ep = L.GemmEpilogue(bias=p(bvec), aux=p(aux), ldaux=N, aux_mode=aux_mode, relu=1 if drop else 0, round_tf32=1 if (mode != 2 and aux_mode != 1) else 0, k_splits=splits, dropout_p=drop, seed=7)
    ts = []
    for i in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.check(lib.st_gemm(mode, p(A), A.shape[1], p(B), B.shape[1], p(Cm), N, M, N, K, C.byref(ep), None)); e1.record()
        torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    t = min(ts); fl = 2.0 * M * N * K
    print(f"mode{mode} M{M:6d} N{N:5d} K{K:6d} splits{splits:3d} aux{aux_mode} drop{drop}: {t*1e3:8.1f} us  {fl / t / 1e9:7.1f} TFLOP/s", flush=True)
if len(sys.argv) > 2 and sys.argv[1] == "--one":
    M, N, K, mode = map(int, sys.argv[2:6]); run(mode, M, N, K, splits=int(sys.argv[6]) if len(sys.argv) > 6 else 1, iters=1)
else:
    for (mode, M, N, K, s) in [(0, 32000, 1536, 512, 1), (0, 32000, 512, 512, 1), (0, 32000, 2048, 512, 1), (0, 32000, 512, 2048, 1),
                               (0, 32000, 1024, 512, 1), (1, 32000, 512, 1536, 1), (1, 32000, 2048, 512, 1), (1, 32000, 512, 2048, 1),
                               (1, 32000, 512, 512, 1), (2, 512, 512, 32000, 37), (2, 2048, 512, 32000, 9), (2, 512, 2048, 32000, 9),
                               (0, 8192, 8192, 8192, 1), (0, 1600, 512, 512, 1), (0, 1600, 2048, 512, 1)]:
        run(mode, M, N, K, s)
    run(0, 32000, 512, 512, aux_mode=1); run(0, 32000, 512, 2048, aux_mode=1); run(0, 32000, 2048, 512, drop=0.1)
    run(1, 32000, 2048, 512, aux_mode=2); run(1, 32000, 512, 2048, aux_mode=1); run(1, 32000, 512, 1536, aux_mode=1)
