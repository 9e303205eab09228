import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
F = stb.functional; L = stb._lib; lib = L.load()
DEV = "cuda:0"
def p(t): return None if t is None else t.data_ptr()
def pad64(n): return (n + 63) // 64 * 64
def rel(a, b): return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()
M, d, f = 200, 64, 128
g = torch.Generator().manual_seed(0)
x = torch.randn(M, d, generator=g).to(DEV); dout = torch.randn(M, d, generator=g).to(DEV)
w1 = (torch.randn(f, d, generator=g) * 0.1).to(DEV); b1 = (torch.randn(f, generator=g) * 0.1).to(DEV)
w2 = (torch.randn(d, f, generator=g) * 0.1).to(DEV); b2 = (torch.randn(d, generator=g) * 0.1).to(DEV)
lg = torch.ones(d, device=DEV); lb = torch.zeros(d, device=DEV)
ns = lib.st_ffn_saved_floats(M, d, f, 0); nw = lib.st_ffn_ws_floats(M, d, f)
for poison in (0.0, float("nan"), 1e30):
    saved = torch.full((ns,), poison, device=DEV); ws = torch.full((nw,), poison, device=DEV); out = torch.full((M, d), poison, device=DEV)
    fa = L.FfnArgs(rows=M, d_model=d, d_ff=f, x=p(x), w1=p(w1), b1=p(b1), w2=p(w2), b2=p(b2), ln_g=p(lg), ln_b=p(lb), eps=1e-6,
                   dropout_p=0.0, seed=0, x_is_tf32=0, round_out=0, out=p(out), saved=p(saved), saved_floats=ns, ws=p(ws), ws_floats=nw)
    L.check(lib.st_ffn_fwd(C.byref(fa), None)); torch.cuda.synchronize()
    grads = [torch.full_like(t, poison) for t in (w1, b1, w2, b2, lg, lb)]; dx = torch.full_like(x, poison)
    ba = L.FfnBwdArgs(f=fa, dout=p(dout), dx=p(dx), dw1=p(grads[0]), db1=p(grads[1]), dw2=p(grads[2]), db2=p(grads[3]), dln_g=p(grads[4]), dln_b=p(grads[5]))
    L.check(lib.st_ffn_bwd(C.byref(ba), None)); torch.cuda.synchronize()
    dz = ws[:M * d].view(M, d); dh = ws[pad64(M * d):pad64(M * d) + M * f].view(M, f)
    if poison == 0.0: ref = [t.clone() for t in (out, dx, dz, dh, *grads)]
    else:
        names = ["out", "dx", "dz", "dh", "dw1", "db1", "dw2", "db2", "dg", "db"]
        for n, a, b in zip(names, (out, dx, dz, dh, *grads), ref):
            bad = ~torch.isclose(a, b, rtol=1e-5, atol=1e-6, equal_nan=False)
            print(f"poison {poison}: {n}: mismatches {int(bad.sum())} of {a.numel()}", (bad.nonzero()[:4].tolist() if bad.any() else ""), flush=True)
