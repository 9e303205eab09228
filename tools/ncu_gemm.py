"""One GEMM launch inside a cudaProfiler range, for `ncu --profile-from-start off`.
python tools/ncu_gemm.py M N K mode [bn] [cluster]"""
import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
L = stb._lib; lib = L.load(); DEV = "cuda:0"
M, N, K, mode = map(int, sys.argv[1:5])
if len(sys.argv) > 5: lib.st_set_option(b"gemm_bn", int(sys.argv[5]))
if len(sys.argv) > 6: lib.st_set_option(b"gemm_cluster", int(sys.argv[6]))
A = torch.randn((K, M) if mode == 2 else (M, K), device=DEV); B = torch.randn((N, K) if mode == 0 else (K, N), device=DEV)
Cm = torch.zeros(M, N, device=DEV)
ep = L.GemmEpilogue(bias=None, aux=None, ldaux=N, aux_mode=0, relu=0, round_tf32=1, k_splits=1, dropout_p=0.0, seed=7)
def go(): L.check(lib.st_gemm(mode, A.data_ptr(), A.shape[1], B.data_ptr(), B.shape[1], Cm.data_ptr(), N, M, N, K, C.byref(ep), None))
go(); torch.cuda.synchronize()
torch.cuda.profiler.start(); go(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
