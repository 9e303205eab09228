"""Ad-hoc per-output error report for the composite operators (GPU box)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import speech_tranformer_pytorch_b200 as stb
from oracle import st_oracle as O
from helpers import relerr
F = stb.functional
DEV = "cuda:0"

def rnd(mod, gen):
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if p.dim() >= 2: torch.nn.init.xavier_normal_(p, generator=gen)
            elif n.endswith("layernorm.weight"): p.copy_(1 + 0.1 * torch.randn(p.shape, generator=gen))
            else: p.copy_(0.05 * torch.randn(p.shape, generator=gen))

def ffn(B, L, d, dff, clean):
    gen = torch.Generator().manual_seed(L)
    m = stb.PositionwiseFeedForward(d, dff).eval(); rnd(m, gen)
    P = {k: v.detach().clone().double().requires_grad_() for k, v in m.state_dict().items()}
    x = torch.randn(B, L, d, generator=gen); g = torch.randn(B, L, d, generator=gen)
    rx = x.clone().double().requires_grad_(); ry = O.positionwise_ffn(rx, P); ry.backward(g.double())
    m = m.to(DEV); cx = x.to(DEV)
    if clean: cx = F.round_tf32(cx)
    cx.requires_grad_()
    if clean: F.mark_tf32_clean(cx)
    cy = m(cx); cy.backward(g.to(DEV))
    errs = {"y": relerr(cy, ry), "dx": relerr(cx.grad, rx.grad)}
    for k, p in m.named_parameters(): errs[k] = relerr(p.grad, P[k].grad)
    print(f"FFN B{B} L{L} d{d} dff{dff} clean={clean}: " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()), flush=True)

def mha(B, L, d, H, masked=True):
    gen = torch.Generator().manual_seed(L + 1)
    m = stb.MultiHeadAttention(H, d, d // H, d // H, return_attention=True).eval(); rnd(m, gen)
    P = {k: v.detach().clone().double().requires_grad_() for k, v in m.state_dict().items()}
    x = torch.randn(B, L, d, generator=gen); g = torch.randn(B, L, d, generator=gen)
    lens = torch.randint(L // 2, L + 1, (B,), generator=gen); lens[0] = L
    mask = O.padding_info_mask(lens, lens).bool() if masked else None
    rx = x.clone().double().requires_grad_(); ry, rw = O.multi_head_attention(rx, rx, rx, mask, P, H); ry.backward(g.double())
    m = m.to(DEV); cx = x.to(DEV).requires_grad_()
    cy, cw = m(cx, cx, cx, None if mask is None else mask.to(DEV)); cy.backward(g.to(DEV))
    errs = {"y": relerr(cy, ry), "attn": relerr(cw, rw), "dx": relerr(cx.grad, rx.grad)}
    sc = max(p.grad.abs().max().item() for p in P.values())
    for k, p in m.named_parameters(): errs[k] = (p.grad.cpu().double() - P[k].grad).abs().max().item() / sc
    print(f"MHA B{B} L{L} d{d} H{H}: " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()), flush=True)

for clean in (False, True):
    ffn(2, 9, 64, 128, clean); ffn(2, 100, 64, 128, clean); ffn(2, 300, 512, 2048, clean)
mha(2, 11, 64, 2); mha(3, 77, 64, 2); mha(2, 300, 512, 8); mha(2, 150, 512, 4); mha(2, 300, 512, 8, masked=False)
