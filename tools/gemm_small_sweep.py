"""Decoder-side GEMM shapes (M = B*L = 1600) under forced tile widths: which BN should pick_bn choose?
python tools/gemm_small_sweep.py [fp16]        (default: TF32 operands)"""
import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
L = stb._lib; lib = L.load(); DEV = "cuda:0"
def p(t): return None if t is None else t.data_ptr()
H16 = len(sys.argv) > 1 and sys.argv[1] == "fp16"
def run(mode, M, N, K, bn):
    lib.st_set_option(b"gemm_bn", bn)
    A = torch.randn((K, M) if mode == 2 else (M, K), device=DEV); B = torch.randn((N, K) if mode == 0 else (K, N), device=DEV)
    Cm = torch.zeros(M, N, device=DEV)
    if H16: A, B, Cm = A.half(), B.half(), Cm.half()
    ep = L.GemmEpilogue(bias=None, aux=None, ldaux=N, aux_mode=0, relu=0, round_tf32=1, k_splits=1, dropout_p=0.0, seed=0)
    def go():
        if H16: L.check(lib.st_gemm_dt(L.DTYPE_F16, mode, p(A), A.shape[1], p(B), B.shape[1], p(Cm), N, 1, M, N, K, C.byref(ep), None))
        else: L.check(lib.st_gemm(mode, p(A), A.shape[1], p(B), B.shape[1], p(Cm), N, M, N, K, C.byref(ep), None))
    for _ in range(3): go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): go()
    e1.record(); torch.cuda.synchronize()
    lib.st_set_option(b"gemm_bn", 0)
    return e0.elapsed_time(e1) / 20 * 1e3
for mode, M, N, K in [(0, 1600, 512, 512), (0, 1600, 512, 2048), (0, 1600, 2048, 512), (0, 1600, 1536, 512), (1, 1600, 512, 512),
                      (1, 1600, 512, 2048), (1, 1600, 2048, 512), (1, 1600, 512, 1536), (0, 320, 512, 512), (0, 320, 1536, 512), (0, 320, 2048, 512), (0, 320, 512, 2048)]:
    print(f"mode{mode} M{M} N{N} K{K}: " + "  ".join(f"bn{bn}: {run(mode, M, N, K, bn):6.1f} us" for bn in (0, 64, 128, 256)), flush=True)
