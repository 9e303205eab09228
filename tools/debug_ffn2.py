import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import speech_tranformer_pytorch_b200 as stb
from oracle import st_oracle as O
F = stb.functional; L = stb._lib; lib = L.load()
DEV = "cuda:0"
def p(t): return None if t is None else t.data_ptr()
def bad(a, b, name):
    a = a.detach().cpu().double(); b = b.detach().cpu().double()
    e = (a - b).abs(); thr = 2e-3 * b.abs().max()
    a2 = e.reshape(-1, e.shape[-1])
    rows = (a2.max(1).values > thr).nonzero().flatten(); cols = (a2.max(0).values > thr).nonzero().flatten()
    print(f"  {name}: relerr {e.max() / b.abs().max():.2e} bad rows {len(rows)} {rows[:8].tolist()}..{rows[-4:].tolist()} bad cols {len(cols)} {cols[:8].tolist()}", flush=True)

def run(B, Lx, d, dff, via):
    gen = torch.Generator().manual_seed(Lx)
    m = stb.PositionwiseFeedForward(d, dff).eval()
    with torch.no_grad():
        for n, q in m.named_parameters():
            if q.dim() >= 2: torch.nn.init.xavier_normal_(q, generator=gen)
            elif n.endswith("layernorm.weight"): q.copy_(1 + 0.1 * torch.randn(q.shape, generator=gen))
            else: q.copy_(0.05 * torch.randn(q.shape, generator=gen))
    P = {k: v.detach().clone().double().requires_grad_() for k, v in m.state_dict().items()}
    x = torch.randn(B, Lx, d, generator=gen); g = torch.randn(B, Lx, d, generator=gen)
    rx = x.clone().double().requires_grad_(); ry = O.positionwise_ffn(rx, P); ry.backward(g.double())
    m = m.to(DEV)
    print(f"FFN B{B} L{Lx} d{d} dff{dff} via {via}")
    if via == "module":
        cx = x.to(DEV).requires_grad_(); cy = m(cx); cy.backward(g.to(DEV))
        bad(cy, ry, "y"); bad(cx.grad, rx.grad, "dx")
        for k, q in m.named_parameters(): bad(q.grad, P[k].grad, k)
    else:
        M = B * Lx
        xs = x.to(DEV).view(M, d).contiguous(); dout = g.to(DEV).view(M, d).contiguous()
        sd = {k: v.detach() for k, v in m.named_parameters()}
        ns = lib.st_ffn_saved_floats(M, d, dff, 0); nw = lib.st_ffn_ws_floats(M, d, dff)
        saved = torch.empty(ns, device=DEV); ws = torch.empty(nw, device=DEV); out = torch.empty(M, d, device=DEV)
        fa = L.FfnArgs(rows=M, d_model=d, d_ff=dff, x=p(xs), w1=p(sd["fc1.weight"]), b1=p(sd["fc1.bias"]), w2=p(sd["fc2.weight"]), b2=p(sd["fc2.bias"]),
                       ln_g=p(sd["layernorm.weight"]), ln_b=p(sd["layernorm.bias"]), eps=1e-6, dropout_p=0.0, seed=0, x_is_tf32=0, round_out=1,
                       out=p(out), saved=p(saved), saved_floats=ns, ws=p(ws), ws_floats=nw)
        L.check(lib.st_ffn_fwd(C.byref(fa), None))
        names = ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "layernorm.weight", "layernorm.bias"]
        grads = [torch.empty_like(sd[n]) for n in names]; dx = torch.empty_like(xs)
        ba = L.FfnBwdArgs(f=fa, dout=p(dout), dx=p(dx), dw1=p(grads[0]), db1=p(grads[1]), dw2=p(grads[2]), db2=p(grads[3]), dln_g=p(grads[4]), dln_b=p(grads[5]))
        L.check(lib.st_ffn_bwd(C.byref(ba), None)); torch.cuda.synchronize()
        bad(out.view(B, Lx, d), ry, "y"); bad(dx.view(B, Lx, d), rx.grad, "dx")
        for n, gq in zip(names, grads): bad(gq, P[n].grad, n)

run(2, 100, 64, 128, "direct")
run(2, 100, 64, 128, "module")
run(2, 100, 64, 128, "direct")
