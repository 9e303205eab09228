"""Time the 16-bit tcgen05 GEMM on the headline shapes with the epilogues the composites use (CUDA events, L2 flushed).
python tools/bench_gemm16.py [bf16|fp16] [--ncu CASE]      CASE = index into the list below (profiler range around one launch)"""
import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
L = stb._lib; lib = L.load(); DEV = "cuda:0"
dt_name = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] in ("bf16", "fp16") else "bf16"
tdt = torch.bfloat16 if dt_name == "bf16" else torch.float16
DT = L.DTYPE_BF16 if dt_name == "bf16" else L.DTYPE_F16
def p(t): return None if t is None else t.data_ptr()
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)
# (name, mode, M, N, K, bias, aux_mode, relu/drop, c_lp, splits)
CASES = [("qkv proj", 0, 32000, 1536, 512, True, 0, 0.0, 1, 1), ("out proj -> z", 0, 32000, 512, 512, True, 1, 0.0, 0, 1),
         ("fc1 relu drop", 0, 32000, 2048, 512, True, 0, 0.1, 1, 1), ("fc2 -> z", 0, 32000, 512, 2048, True, 1, 0.0, 0, 1),
         ("dh = dz W2 mask", 1, 32000, 2048, 512, False, 2, 0.0, 1, 1), ("dx = dh W1 + dz", 1, 32000, 512, 2048, False, 1, 0.0, 1, 1),
         ("dx = dqkv W + dz", 1, 32000, 512, 1536, False, 1, 0.0, 1, 1), ("dctx = dz Wo", 1, 32000, 512, 512, False, 0, 0.0, 1, 1),
         ("dW1", 2, 2048, 512, 32000, False, 0, 0.0, 0, 9), ("dW2", 2, 512, 2048, 32000, False, 0, 0.0, 0, 9),
         ("dWqkv", 2, 1536, 512, 32000, False, 0, 0.0, 0, 12), ("dWo", 2, 512, 512, 32000, False, 0, 0.0, 0, 37),
         ("square 8192", 0, 8192, 8192, 8192, False, 0, 0.0, 1, 1)]
def make(case):
    name, mode, M, N, K, bias, aux_mode, drop, c_lp, splits = case
    A = torch.randn((K, M) if mode == 2 else (M, K), device=DEV).to(tdt); B = torch.randn((N, K) if mode == 0 else (K, N), device=DEV).to(tdt)
    Cm = torch.zeros(M, N, device=DEV, dtype=tdt if c_lp else torch.float32)
    bvec = torch.randn(N, device=DEV) if bias else None
    aux = torch.randn(M, N, device=DEV).to(tdt) if aux_mode else None
    colsum = None
    ep = L.GemmEpilogue(bias=p(bvec), aux=p(aux), ldaux=N, aux_mode=aux_mode, relu=1 if drop else 0, round_tf32=0, k_splits=splits, dropout_p=drop, seed=7)
    def go(): L.check(lib.st_gemm_dt(DT, mode, p(A), A.shape[1], p(B), B.shape[1], p(Cm), N, c_lp, M, N, K, C.byref(ep), None))
    return go, (A, B, Cm, bvec, aux)
if "--ncu" in sys.argv:
    case = CASES[int(sys.argv[sys.argv.index("--ncu") + 1])]
    go, keep = make(case); go(); torch.cuda.synchronize()
    torch.cuda.profiler.start(); go(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
    sys.exit(0)
for case in CASES:
    name, mode, M, N, K, bias, aux_mode, drop, c_lp, splits = case
    go, keep = make(case)
    ts = []
    for i in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    t = min(ts); fl = 2.0 * M * N * K
    print(f"{dt_name} {name:20s} mode{mode} M{M:6d} N{N:5d} K{K:6d}: {t*1e3:8.1f} us  {fl / t / 1e9:7.1f} TFLOP/s", flush=True)
    del go, keep
