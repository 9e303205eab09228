import sys, os, ctypes as C, torch
sys.path.insert(0, "/root/repo")
import speech_tranformer_pytorch_b200 as stb
L = stb._lib; lib = L.load(); DEV = "cuda:0"
def p(t): return None if t is None else t.data_ptr()
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)
ok = True
for tdt, DT in ((torch.float16, L.DTYPE_F16), (torch.bfloat16, L.DTYPE_BF16)):
    for (mode, M, N, K, bias, drop) in [(0, 32000, 1536, 512, True, 0.0), (0, 32000, 2048, 512, True, 0.1), (1, 32000, 512, 512, False, 0.0),
                                        (0, 1600, 1536, 512, True, 0.0), (0, 333, 520, 192, True, 0.1), (0, 32000, 1024, 512, True, 0.0), (1, 77, 72, 64, False, 0.0)]:
        torch.manual_seed(1)
        A = torch.randn(M, K, device=DEV).to(tdt); B = torch.randn((N, K) if mode == 0 else (K, N), device=DEV).to(tdt)
        bvec = torch.randn(N, device=DEV) if bias else None
        ep = L.GemmEpilogue(bias=p(bvec), aux=None, ldaux=0, aux_mode=0, relu=1 if drop else 0, round_tf32=0, k_splits=1, dropout_p=drop, seed=7)
        outs, times = [], []
        for opt in (0, 1, 2):
            lib.st_set_option(b"gemm_rows16", opt)
            Cm = torch.zeros(M, N, device=DEV, dtype=tdt)
            go = lambda: L.check(lib.st_gemm_dt(DT, mode, p(A), K, p(B), B.shape[1], p(Cm), N, 1, M, N, K, C.byref(ep), None))
            go(); torch.cuda.synchronize()
            ts = []
            for i in range(7):
                flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); go(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
            outs.append(Cm.clone()); times.append(sorted(ts)[len(ts) // 2])
        same = torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
        ok &= same
        print(f"{str(tdt)[6:]:9s} mode{mode} M{M} N{N} K{K} drop{drop}: identical={same}  us transposed/direct/staged = " + " / ".join(f"{t:.1f}" for t in times))
lib.st_set_option(b"gemm_rows16", 0)
print("ALL IDENTICAL" if ok else "MISMATCH")
