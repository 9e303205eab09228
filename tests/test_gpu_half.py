"""GPU parity of the 16-bit path (fp16 / bf16 activations and tensor-core operands, fp32 accumulation, statistics,
parameters and parameter gradients; BASELINE.json configs[2]) and of length-aware masking, through the C ABI, against
the float64 CPU oracle (oracle/st_oracle.py).

Tolerances (metric max|a-b| / max|b|, per tensor):
  fp16: operands carry the same 10-bit mantissa as TF32, so the module bound stays the north-star 1e-3 for gradients;
        OUTPUTS are additionally rounded to fp16 when stored (<= 2^-11 = 4.9e-4 of the largest value), hence 1.5e-3 there.
  bf16: 8-bit mantissa — every operand rounding is 2^-9 = 2e-3 and the stored output another 2e-3 of its value; bound 2e-2
        (measured values are written to gpurun_out/half_parity_errors.json by the last test of this file).
The oracle is fed the SAME 16-bit-representable inputs the module receives and the fp32 master parameters."""
import json
import os

import pytest
import torch

from helpers import TOL, relerr, relu_gate_from_cuda
from oracle import st_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
DTYPES = [torch.float16, torch.bfloat16]
TOL_OUT = {torch.float16: 1.5e-3, torch.bfloat16: 2e-2}
TOL_GRAD = {torch.float16: 1.5e-3, torch.bfloat16: 2e-2}
TOL_ATTN16 = {torch.float16: 2.5e-3, torch.bfloat16: 2e-2}   # raw attention core on N(0,1) q/k/v (cf. TOL_ATTN in test_gpu_parity)
MEASURED = {}


def _rec(name, dtype, **errs):
    MEASURED[f"{name}[{str(dtype).split('.')[-1]}]"] = {k: float(v) for k, v in errs.items()}


@pytest.fixture(scope="module")
def stb():
    import speech_tranformer_pytorch_b200 as m
    m.build()
    m._lib.check(m._lib.load().st_device_check(0))
    return m


def _grad_check(cuda_params, oracle_params, tol, zero_ok=("linear_k.bias",)):
    """Per-tensor metric; a parameter whose true gradient is analytically zero (the key bias: softmax is shift
    invariant) is compared against the largest gradient of the module instead."""
    worst = 0.0
    scale = max(p.grad.abs().max().item() for p in oracle_params.values())
    for k, p in oracle_params.items():
        got = cuda_params[k].grad
        assert got is not None, k
        if any(k.endswith(z) for z in zero_ok):
            err = (got.detach().cpu().double() - p.grad.double()).abs().max().item() / scale
        else:
            err = relerr(got, p.grad)
        assert err < tol, (k, err)
        worst = max(worst, err)
    return worst


def _ref_attn(q, k, v, mask, H):
    """Attention.py:78-90 through the oracle's single-head routine (heads folded into the batch)."""
    B, Lq, d = q.shape
    Lk, dk = k.shape[1], d // H
    sh = lambda x: x.view(B, -1, H, dk).transpose(1, 2).reshape(B * H, -1, dk)
    m = None if mask is None else mask.unsqueeze(1).expand(B, H, Lq, Lk).reshape(B * H, Lq, Lk)
    o, w = O.scaled_dot_product_attention(sh(q), sh(k), sh(v), m, dk)
    return o.view(B, H, Lq, dk).transpose(1, 2).reshape(B, Lq, d), w.view(B, H, Lq, Lk)


# ------------------------------------------------------------------------------------------------ attention core
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,H,Lq,Lk,kind", [(2, 2, 128, 128, "none"), (2, 8, 300, 300, "lengths"), (3, 2, 50, 50, "causal"),
                                            (2, 4, 50, 333, "lengths"), (2, 2, 77, 200, "dense"), (1, 1, 1, 1, "none"),
                                            (2, 8, 1000, 1000, "lengths")])
def test_attention_core_16bit(stb, dtype, B, H, Lq, Lk, kind):
    F = stb.functional
    dk = 64
    d = H * dk
    gen = torch.Generator().manual_seed(B * 1000 + Lq + Lk)
    q = torch.randn(B, Lq, d, generator=gen).to(dtype)
    k = torch.randn(B, Lk, d, generator=gen).to(dtype)
    v = torch.randn(B, Lk, d, generator=gen).to(dtype)
    g = torch.randn(B, Lq, d, generator=gen).to(dtype)
    lens = torch.randint(max(1, Lk // 3), Lk + 1, (B,), generator=gen)
    lens[0] = Lk
    if kind == "none":
        mask_c, mask_o = None, None
    elif kind == "lengths":
        mask_c = F.LengthMask(lens.to(DEV), Lq, Lk)
        mask_o = O.padding_info_mask(torch.full((B,), Lq), lens).bool()
    elif kind == "causal":
        mask_c = F.LengthMask(lens.to(DEV), Lq, Lk, causal=True)
        mask_o = O.padding_info_mask(lens, lens).bool() | O.feature_info_mask(lens).bool()
    else:
        mask_o = torch.rand(B, Lq, Lk, generator=gen) < 0.3
        mask_o[:, :, 0] = False            # no fully masked row
        mask_c = mask_o.to(DEV)
    if mask_c is not None and kind != "dense":   # the length-derived mask is bit-identical to the reference's mask tensor
        assert torch.equal(mask_c.dense().cpu(), mask_o)
    cq, ck, cv = (x.to(DEV).requires_grad_() for x in (q, k, v))
    out, attn = F.attention_core(cq, ck, cv, mask_c, n_head=H, need_attn=True)
    assert out.dtype == dtype
    out.backward(g.to(DEV))
    rq, rk, rv = (x.double().requires_grad_() for x in (q, k, v))
    ro, rw = _ref_attn(rq, rk, rv, mask_o, H)
    ro.backward(g.double())
    errs = dict(out=relerr(out, ro), attn=relerr(attn, rw), dq=relerr(cq.grad, rq.grad), dk=relerr(ck.grad, rk.grad),
                dv=relerr(cv.grad, rv.grad))
    _rec(f"attention_core B{B}H{H}Lq{Lq}Lk{Lk}{kind}", dtype, **errs)
    assert max(errs.values()) < TOL_ATTN16[dtype], errs
    if mask_o is not None:                 # masked probabilities are exactly zero
        assert torch.all(attn.cpu()[mask_o.unsqueeze(1).expand(-1, H, -1, -1)] == 0)


@pytest.mark.parametrize("dtype", [torch.float32] + DTYPES)
def test_length_mask_equals_dense_mask_bitwise(stb, dtype):
    """Lengths + causal flag give bit-identical results to the mask tensors Utils.py:41-70 builds — forward and backward."""
    F = stb.functional
    B, H, L, d = 3, 2, 150, 128
    gen = torch.Generator().manual_seed(5)
    lens = torch.tensor([150, 97, 31])
    x = torch.randn(B, L, d, generator=gen).to(dtype)
    g = torch.randn(B, L, d, generator=gen).to(dtype)
    for causal in (False, True):
        dense = O.padding_info_mask(lens, lens).bool()
        if causal:
            dense = dense | O.feature_info_mask(lens).bool()
        res = []
        for m in (F.LengthMask(lens.to(DEV), L, L, causal=causal), dense.to(DEV)):
            cx = x.to(DEV).requires_grad_()
            out, attn = F.attention_core(cx, cx, cx, m, n_head=H, need_attn=True)
            out.backward(g.to(DEV))
            res.append((out, attn, cx.grad))
        for a, b in zip(*res):
            assert torch.equal(a, b)


@pytest.mark.parametrize("dtype", DTYPES)
def test_attention_dropout_16bit_forward_backward_consistent(stb, dtype):
    """The backward kernels regenerate the forward dropout mask: with V = identity-like probes the gradient of a kept
    probability is non-zero exactly where the returned (post-dropout) weights are non-zero."""
    F = stb.functional
    B, H, L, d = 1, 1, 128, 64
    gen = torch.Generator().manual_seed(3)
    q = torch.randn(B, L, d, generator=gen).to(dtype).to(DEV)
    out1, w1 = F.attention_core(q, q, q, None, n_head=H, dropout_p=0.3, seed=1234, need_attn=True)
    out2, w2 = F.attention_core(q, q, q, None, n_head=H, dropout_p=0.3, seed=1234, need_attn=True)
    assert torch.equal(out1, out2) and torch.equal(w1, w2)
    kept = (w1 != 0).float().mean().item()
    assert 0.6 < kept < 0.8
    # out == W_dropped @ v up to 16-bit rounding: the P·V MMA used the same mask the returned weights show
    ref = (w1[0, 0].double() @ q[0].double())
    assert relerr(out1[0], ref) < TOL_ATTN16[dtype]
    # backward: d out / d v through the dropped weights
    v = q.clone().requires_grad_()
    o, w = F.attention_core(q, q, v, None, n_head=H, dropout_p=0.3, seed=77, need_attn=True)
    gsel = torch.randn(B, L, d, generator=gen).to(dtype).to(DEV)
    o.backward(gsel)
    assert relerr(v.grad[0], w[0, 0].double().t() @ gsel[0].double()) < TOL_ATTN16[dtype]


# ------------------------------------------------------------------------------------------------ modules
def _init(m, gen):
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() >= 2:
                torch.nn.init.xavier_normal_(p, generator=gen)
            elif n.endswith("layernorm.weight") or n.endswith("3.weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=gen))
            else:
                p.copy_(0.05 * torch.randn(p.shape, generator=gen))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,Lq,Lk,H,cross", [(2, 200, 200, 8, False), (2, 50, 333, 8, True), (3, 40, 40, 2, False)])
def test_mha_module_16bit(stb, dtype, B, Lq, Lk, H, cross):
    F = stb.functional
    d = 64 * H
    gen = torch.Generator().manual_seed(Lq * 7 + Lk)
    m = stb.MultiHeadAttention(H, d, 64, 64, residual="q" if cross else "v").eval()
    _init(m, gen)
    P = {k: v.detach().clone().double().requires_grad_() for k, v in m.state_dict().items()}
    q = torch.randn(B, Lq, d, generator=gen).to(dtype)
    kv = torch.randn(B, Lk, d, generator=gen).to(dtype) if cross else q
    g = torch.randn(B, Lq, d, generator=gen).to(dtype)
    lens = torch.randint(Lk // 2, Lk + 1, (B,), generator=gen)
    lens[0] = Lk
    causal = (not cross) and Lq == 40
    mask_o = O.padding_info_mask(torch.full((B,), Lq), lens).bool()
    if causal:
        mask_o = mask_o | O.feature_info_mask(lens).bool()
    m = m.to(DEV)
    cq = q.to(DEV).requires_grad_()
    ckv = kv.to(DEV).requires_grad_() if cross else cq
    out, _ = m(cq, ckv, ckv, F.LengthMask(lens.to(DEV), Lq, Lk, causal=causal))
    assert out.dtype == dtype
    out.backward(g.to(DEV))
    rq = q.double().requires_grad_()
    rkv = kv.double().requires_grad_() if cross else rq
    ro, _ = O.multi_head_attention(rq, rkv, rkv, mask_o, P, H, residual="q" if cross else "v")
    ro.backward(g.double())
    e_out, e_dq = relerr(out, ro), relerr(cq.grad, rq.grad)
    e_dkv = relerr(ckv.grad, rkv.grad) if cross else 0.0
    assert e_out < TOL_OUT[dtype] and e_dq < TOL_GRAD[dtype] and e_dkv < TOL_GRAD[dtype], (e_out, e_dq, e_dkv)
    e_p = _grad_check(dict(m.named_parameters()), P, TOL_GRAD[dtype])
    _rec(f"mha B{B}Lq{Lq}Lk{Lk}H{H}", dtype, out=e_out, dq=e_dq, dkv=e_dkv, params=e_p)


@pytest.mark.parametrize("dtype", DTYPES)
def test_ffn_module_16bit(stb, dtype):
    gen = torch.Generator().manual_seed(21)
    m = stb.PositionwiseFeedForward(512, 2048).eval()
    _init(m, gen)
    P = {k: v.detach().clone().double().requires_grad_() for k, v in m.state_dict().items()}
    x = torch.randn(3, 100, 512, generator=gen).to(dtype)
    g = torch.randn(3, 100, 512, generator=gen).to(dtype)
    m = m.to(DEV)
    m.keep_hidden = True
    cx = x.to(DEV).requires_grad_()
    y = m(cx)
    assert y.dtype == dtype and m.last_hidden.dtype == dtype
    y.backward(g.to(DEV))
    rx = x.double().requires_grad_()
    gate = relu_gate_from_cuda(m.last_hidden.float(), O.ffn_preactivation(rx, P))
    ry = O.positionwise_ffn(rx, P, gate=gate)
    ry.backward(g.double())
    e_out, e_dx = relerr(y, ry), relerr(cx.grad, rx.grad)
    assert e_out < TOL_OUT[dtype] and e_dx < TOL_GRAD[dtype], (e_out, e_dx)
    e_p = _grad_check(dict(m.named_parameters()), P, TOL_GRAD[dtype])
    _rec("ffn 300x512x2048", dtype, out=e_out, dx=e_dx, params=e_p)


class _EncoderLayer(torch.nn.Module):
    """Layers.py:8-22 against the drop-in modules."""

    def __init__(self, stb, d_model, d_inner, n_head):
        super().__init__()
        self.slf_attn = stb.MultiHeadAttention(n_head, d_model, d_model // n_head, d_model // n_head)
        self.pos_ffn = stb.PositionwiseFeedForward(d_model, d_inner)

    def forward(self, x, mask=None):
        a, w = self.slf_attn(x, x, x, mask=mask)
        return self.pos_ffn(a), w


@pytest.mark.parametrize("dtype", [torch.float32] + DTYPES)
def test_encoder_layer_headline_length(stb, dtype):
    """EncoderLayer at the headline sequence length and width (B=2, T=1000, d=512, h=8, d_ff=2048; 8 key tiles, every
    pipelined attention-backward CTA walks 16 streamed tiles) against the float64 oracle, ragged lengths."""
    F = stb.functional
    B, L, d, H, dff = 2, 1000, 512, 8, 2048
    gen = torch.Generator().manual_seed(1000)
    m = _EncoderLayer(stb, d, dff, H).eval()
    _init(m, gen)
    P = {k: v.detach().clone().double().requires_grad_() for k, v in m.state_dict().items()}
    x = torch.randn(B, L, d, generator=gen).to(dtype)
    g = torch.randn(B, L, d, generator=gen).to(dtype)
    lens = torch.tensor([L, 641])
    mask = O.padding_info_mask(lens, lens).bool()
    m = m.to(DEV)
    m.pos_ffn.keep_hidden = True
    cx = x.to(DEV).requires_grad_()
    cy, _ = m(cx, F.LengthMask(lens.to(DEV), L, L))
    cy.backward(g.to(DEV))
    rx = x.double().requires_grad_()
    a, _ = O.multi_head_attention(rx, rx, rx, mask, {k[9:]: v for k, v in P.items() if k.startswith("slf_attn.")}, H)
    gate = relu_gate_from_cuda(m.pos_ffn.last_hidden.float(), O.ffn_preactivation(a, {k[8:]: v for k, v in P.items() if k.startswith("pos_ffn.")}))
    ry = O.encoder_layer(rx, mask, P, H, ffn_gate=gate)
    ry.backward(g.double())
    tol_o = TOL if dtype == torch.float32 else TOL_OUT[dtype]
    tol_g = TOL if dtype == torch.float32 else TOL_GRAD[dtype]
    e_out, e_dx = relerr(cy, ry), relerr(cx.grad, rx.grad)
    assert e_out < tol_o and e_dx < tol_g, (e_out, e_dx)
    e_p = _grad_check(dict(m.named_parameters()), P, tol_g)
    _rec("encoder_layer B2T1000d512h8", dtype, out=e_out, dx=e_dx, params=e_p)


@pytest.mark.parametrize("dtype", DTYPES)
def test_error_behaviour_16bit(stb, dtype):
    F = stb.functional
    x = torch.randn(2, 16, 64, device=DEV).to(dtype)
    with pytest.raises(RuntimeError):       # d_k = 32 has no 16-bit kernel
        F.attention_core(x, x, x, None, n_head=2)
    with pytest.raises(RuntimeError):       # mixed activation types
        F.attention_core(x, x.float(), x, None, n_head=1)
    with pytest.raises(RuntimeError):
        F.attention_core(x.cpu(), x.cpu(), x.cpu(), None, n_head=1)


def test_write_measured_errors():
    """Not a check: persists the errors measured above next to the run (gpurun_out/ travels back from the GPU box)."""
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "half_parity_errors.json"), "w") as f:
        json.dump(MEASURED, f, indent=1, sort_keys=True)
