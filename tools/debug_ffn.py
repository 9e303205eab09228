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
MODE = sys.argv[1] if len(sys.argv) > 1 else "a"
w1 = (torch.randn(f, d, generator=g) * 0.1).to(DEV); b1 = (torch.randn(f, generator=g) * 0.1).to(DEV)
w2 = (torch.randn(d, f, generator=g) * 0.1).to(DEV); b2 = (torch.randn(d, generator=g) * 0.1).to(DEV)
lg = torch.ones(d, device=DEV); lb = torch.zeros(d, device=DEV)
if MODE in ("b", "c"):
    lg = (1 + 0.1 * torch.randn(d, generator=g)).to(DEV); lb = (0.05 * torch.randn(d, generator=g)).to(DEV)
if MODE == "c":
    w1 = torch.nn.init.xavier_normal_(torch.empty(f, d), generator=g).to(DEV); w2 = torch.nn.init.xavier_normal_(torch.empty(d, f), generator=g).to(DEV)
print("MODE", MODE)
ns = lib.st_ffn_saved_floats(M, d, f, 0); nw = lib.st_ffn_ws_floats(M, d, f)
saved = torch.zeros(ns, device=DEV); ws = torch.zeros(nw, device=DEV); out = torch.empty(M, d, device=DEV)
fa = L.FfnArgs(rows=M, d_model=d, d_ff=f, x=p(x), w1=p(w1), b1=p(b1), w2=p(w2), b2=p(b2), ln_g=p(lg), ln_b=p(lb), eps=1e-6,
               dropout_p=0.0, seed=0, x_is_tf32=0, round_out=0, out=p(out), saved=p(saved), saved_floats=ns, ws=p(ws), ws_floats=nw)
L.check(lib.st_ffn_fwd(C.byref(fa), None)); torch.cuda.synchronize()
off = 0
xr = saved[off:off + M * d].view(M, d); off += pad64(M * d)
H = saved[off:off + M * f].view(M, f); off += pad64(M * f)
z = saved[off:off + M * d].view(M, d); off += pad64(M * d)
mean = saved[off:off + M]; off += pad64(M); rstd = saved[off:off + M]; off += pad64(M)
w1r = saved[off:off + f * d].view(f, d); off += pad64(f * d); w2r = saved[off:off + f * d].view(d, f)
Href = torch.relu(xr.double() @ w1r.double().t() + b1.double())
print("xr", rel(xr, x), "H", rel(H, Href), "z", rel(z, H.double() @ w2r.double().t() + b2.double() + x.double()), "w2r", rel(w2r, w2))
grads = [torch.empty_like(t) for t in (w1, b1, w2, b2, lg, lb)]; dx = torch.empty_like(x)
ba = L.FfnBwdArgs(f=fa, dout=p(dout), dx=p(dx), dw1=p(grads[0]), db1=p(grads[1]), dw2=p(grads[2]), db2=p(grads[3]), dln_g=p(grads[4]), dln_b=p(grads[5]))
L.check(lib.st_ffn_bwd(C.byref(ba), None)); torch.cuda.synchronize()
dz = ws[:M * d].view(M, d); dh = ws[pad64(M * d):pad64(M * d) + M * f].view(M, f)
zz = z.double().requires_grad_(); y = torch.nn.functional.layer_norm(zz, (d,), lg.double(), lb.double(), 1e-6); y.backward(dout.double())
print("dz", rel(dz, zz.grad))
dh_ref = (dz.double() @ w2r.double()) * (H > 0)
e = (dh.double() - dh_ref).abs()
print("dh", rel(dh, dh_ref), "bad rows", (e.max(1).values > 1e-3).nonzero().flatten()[:10].tolist(), "bad cols", (e.max(0).values > 1e-3).nonzero().flatten()[:10].tolist())
print("db1", rel(grads[1], dh_ref.sum(0)), "db1(from dh)", rel(grads[1], dh.double().sum(0)))
print("dw1", rel(grads[0], dh_ref.t() @ xr.double()), "dw1(from dh)", rel(grads[0], dh.double().t() @ xr.double()))
print("dx", rel(dx, dh_ref @ w1r.double() + dz.double()))
print("dw2", rel(grads[2], dz.double().t() @ H.double()))
