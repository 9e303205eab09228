"""Randomised shape sweep: EncoderLayer and DecoderLayer (self + cross attention, FFN) forward / backward against the
float64 oracle over random (B, L, T, d_model, heads, d_ff, lengths).  python tests/fuzz_layers.py [n_cases] [seed]"""
import os, sys, random, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import speech_tranformer_pytorch_b200 as stb
from speech_tranformer_pytorch_b200 import model as smodel
from helpers import relerr, relu_gate_from_cuda
from oracle import st_oracle as O
DEV = "cuda:0"
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = 0.0
for case in range(n_cases):
    dk = rng.choice([32, 64, 128]); H = rng.choice([1, 2, 4, 8]) if dk < 128 else rng.choice([1, 2, 4])
    d = dk * H; dff = rng.choice([128, 384, 1024, 2048]); B = rng.randint(1, 4)
    T = rng.choice([1, 7, 63, 64, 65, 129, 200, 333, 520]); L = rng.choice([1, 3, 17, 50, 64, 65, 130])
    gen = torch.Generator().manual_seed(case)
    layer = smodel.DecoderLayer(d, dff, H, dk, dk).eval()
    enc_layer = smodel.EncoderLayer(d, dff, H, dk, dk).eval()
    for m in (layer, enc_layer):
        with torch.no_grad():
            for n, p in m.named_parameters():
                if p.dim() >= 2: torch.nn.init.xavier_normal_(p, generator=gen)
                elif n.endswith("layernorm.weight"): p.copy_(1 + 0.1 * torch.randn(p.shape, generator=gen))
                else: p.copy_(0.05 * torch.randn(p.shape, generator=gen))
    Pd = {k: v.detach().clone().double().requires_grad_() for k, v in layer.state_dict().items()}
    Pe = {k: v.detach().clone().double().requires_grad_() for k, v in enc_layer.state_dict().items()}
    x = torch.randn(B, T, d, generator=gen); y = torch.randn(B, L, d, generator=gen)
    gx = torch.randn(B, T, d, generator=gen); gy = torch.randn(B, L, d, generator=gen)
    tl = torch.randint(1, T + 1, (B,), generator=gen); tl[0] = T
    ll = torch.randint(1, L + 1, (B,), generator=gen); ll[0] = L
    enc_mask = O.padding_info_mask(tl, tl).bool(); slf_mask = O.decoder_self_mask(ll); cross_mask = O.padding_info_mask(ll, tl).bool()
    layer, enc_layer = layer.to(DEV), enc_layer.to(DEV)
    layer.pos_ffn.keep_hidden = enc_layer.pos_ffn.keep_hidden = True
    cx = x.to(DEV).requires_grad_(); cy = y.to(DEV).requires_grad_()
    ce, _ = enc_layer(cx, enc_mask.to(DEV))
    cd, _ = layer(cy, ce, slf_mask.to(DEV), cross_mask.to(DEV))
    (ce * gx.to(DEV)).sum().backward(retain_graph=True); cd.backward(gy.to(DEV))
    rx = x.double().requires_grad_(); ry = y.double().requires_grad_()
    sub = lambda P, pre: {k[len(pre):]: v for k, v in P.items() if k.startswith(pre)}
    a, _ = O.multi_head_attention(rx, rx, rx, enc_mask, sub(Pe, "slf_attn."), H)
    re = O.positionwise_ffn(a, sub(Pe, "pos_ffn."), gate=relu_gate_from_cuda(enc_layer.pos_ffn.last_hidden, O.ffn_preactivation(a, sub(Pe, "pos_ffn."))))
    s1, _ = O.multi_head_attention(ry, ry, ry, slf_mask, sub(Pd, "slf_attn."), H)
    c1, _ = O.multi_head_attention(s1, re, re, cross_mask, sub(Pd, "enc_attn."), H, residual="q")
    rd = O.positionwise_ffn(c1, sub(Pd, "pos_ffn."), gate=relu_gate_from_cuda(layer.pos_ffn.last_hidden, O.ffn_preactivation(c1, sub(Pd, "pos_ffn."))))
    (re * gx.double()).sum().backward(retain_graph=True); rd.backward(gy.double())
    errs = {"enc": relerr(ce, re), "dec": relerr(cd, rd), "dx": relerr(cx.grad, rx.grad), "dy": relerr(cy.grad, ry.grad)}
    for (m, P, tag) in ((layer, Pd, "dec."), (enc_layer, Pe, "enc.")):
        scale = max(p.grad.abs().max().item() for p in P.values())
        for k, p in m.named_parameters():
            errs[tag + k] = (p.grad.detach().cpu().double() - P[k].grad).abs().max().item() / scale
    w = max(errs.values()); worst = max(worst, w)
    print(f"case {case:2d} B={B} T={T} L={L} d={d} H={H} dk={dk} dff={dff} lens={tl.tolist()}/{ll.tolist()}: worst {w:.2e} ({max(errs, key=errs.get)})", flush=True)
    assert w < 2e-3, errs
print(f"fuzz ok, worst {worst:.2e}")
