"""CTA-pair GEMM with cluster-launch-control tile scheduling (option gemm_clc) against the static persistent schedule:
bit-identical results and timing on the headline shapes.  python tools/check_clc.py"""
import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
L = stb._lib; lib = L.load(); DEV = "cuda:0"
p = lambda t: None if t is None else t.data_ptr()
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)
ok = True
for tdt, DT in ((torch.float16, L.DTYPE_F16), (torch.float32, L.DTYPE_F32)):
    for (mode, M, N, K, bias, drop, c_lp, aux_mode) in [(0, 32000, 1536, 512, True, 0.0, 1, 0), (0, 32000, 2048, 512, True, 0.1, 1, 0),
                                                        (0, 32000, 512, 2048, True, 0.0, 0, 1), (1, 32000, 512, 1536, False, 0.0, 0, 1),
                                                        (1, 32000, 2048, 512, False, 0.0, 1, 2), (0, 5000, 1024, 256, True, 0.0, 1, 0),
                                                        (0, 64000, 512, 512, True, 0.0, 0, 1)]:
        if tdt == torch.float32: c_lp = 0
        torch.manual_seed(1)
        A = torch.randn(M, K, device=DEV).to(tdt); B = torch.randn((N, K) if mode == 0 else (K, N), device=DEV).to(tdt)
        bvec = torch.randn(N, device=DEV) if bias else None
        aux = torch.randn(M, N, device=DEV).to(tdt) if aux_mode else None
        ep = L.GemmEpilogue(bias=p(bvec), aux=p(aux), ldaux=N, aux_mode=aux_mode, relu=1 if drop else 0, round_tf32=0, k_splits=1, dropout_p=drop, seed=7)
        outs, times = [], []
        for opt in (0, 1):
            lib.st_set_option(b"gemm_clc", opt)
            Cm = torch.zeros(M, N, device=DEV, dtype=tdt if c_lp else torch.float32)
            go = lambda: L.check(lib.st_gemm_dt(DT, mode, p(A), K, p(B), B.shape[1], p(Cm), N, c_lp, M, N, K, C.byref(ep), None))
            go(); torch.cuda.synchronize()
            ts = []
            for i in range(7):
                flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); go(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
            outs.append(Cm.clone()); times.append(sorted(ts)[len(ts) // 2])
        same = torch.equal(outs[0], outs[1]); ok &= same
        print(f"{str(tdt)[6:]:8s} mode{mode} M{M} N{N} K{K} aux{aux_mode} drop{drop}: identical={same}  us static/clc = {times[0]:.1f} / {times[1]:.1f}")
lib.st_set_option(b"gemm_clc", 0)
print("ALL IDENTICAL" if ok else "MISMATCH")
