"""GPU tests of the fp32-boundary / fp16-operand engine of the composite operators (ST_DTYPE_F32_H16, include/st_b200.h;
functional.set_fp32_engine("fp16"), the default for fp32 tensors when d_k = 64).  Parity of this engine against the oracle and
the golden fixtures is covered by the `engine`-parametrised tests of test_gpu_parity.py and by test_gpu_round2.py; here:
what is specific to it — the device-derived power-of-two gradient scale, the boundary types, and the trainer's twins."""
import pytest
import torch

from helpers import TOL, relerr
from oracle import st_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def stb():
    import speech_tranformer_pytorch_b200 as m
    m.build()
    m._lib.check(m._lib.load().st_device_check(0))
    return m


@pytest.fixture(autouse=True)
def _fp16_engine():
    from speech_tranformer_pytorch_b200 import functional as F
    prev = F.set_fp32_engine("fp16")
    yield
    F.set_fp32_engine(prev)


def _layer(stb, d=512, H=8, dff=2048, seed=3, residual="v"):
    gen = torch.Generator().manual_seed(seed)
    att = stb.MultiHeadAttention(H, d, d // H, d // H, residual=residual).eval()
    ffn = stb.PositionwiseFeedForward(d, dff).eval()
    with torch.no_grad():
        for m in (att, ffn):
            for n, p in m.named_parameters():
                if p.dim() >= 2:
                    torch.nn.init.xavier_normal_(p, generator=gen)
                elif n.endswith("layernorm.weight"):
                    p.copy_(1 + 0.1 * torch.randn(p.shape, generator=gen))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=gen))
    return att.to(DEV), ffn.to(DEV)


def _run(att, ffn, x, mask, g):
    for p in list(att.parameters()) + list(ffn.parameters()):
        p.grad = None
    cx = x.clone().requires_grad_()
    y = ffn(att(cx, cx, cx, mask=mask)[0])
    y.backward(g)
    return y.detach(), cx.grad, [p.grad.clone() for p in list(att.parameters()) + list(ffn.parameters())]


@pytest.mark.parametrize("k", [-60, -24, 30])
def test_gradient_scale_is_derived_from_the_incoming_gradient(stb, k):
    """The backward operators scale their 16-bit tensors by a power of two taken from max|grad_output|, so gradients of ANY
    magnitude go through the same fp16 arithmetic: multiplying grad_output by 2^k multiplies every result by exactly 2^k
    (bit for bit) — nothing underflows at 2^-60 and nothing overflows at 2^30."""
    att, ffn = _layer(stb)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 150, 512, generator=gen).to(DEV)
    g = torch.randn(2, 150, 512, generator=gen).to(DEV)
    mask = O.padding_info_mask(torch.tensor([150, 90]), torch.tensor([150, 90])).bool().to(DEV)
    y0, dx0, gr0 = _run(att, ffn, x, mask, g)
    y1, dx1, gr1 = _run(att, ffn, x, mask, g * 2.0 ** k)
    assert torch.equal(y0, y1)
    assert dx0.abs().max().item() > 0
    assert torch.equal(dx1, dx0 * 2.0 ** k)
    for a, b in zip(gr1, gr0):
        # parameter gradients are sums of atomically accumulated partial tiles (split-K, column sums): equal up to the
        # fp32 summation order, which varies from launch to launch; dx above went through every scaled tensor of both
        # operators and is bit-exact
        assert relerr(a, b * 2.0 ** k) < 2e-4


def test_zero_and_nonfinite_incoming_gradient(stb):
    """max|grad_output| = 0 uses scale 1 (all gradients exactly zero); a non-finite grad_output propagates as non-finite
    gradients (so the trainer's non-finite-norm step skip sees it) instead of being masked by the scale."""
    att, ffn = _layer(stb, d=128, H=2, dff=256)
    x = torch.randn(1, 40, 128, generator=torch.Generator().manual_seed(1)).to(DEV)
    _, dx, grads = _run(att, ffn, x, None, torch.zeros(1, 40, 128, device=DEV))
    assert dx.abs().max().item() == 0 and all(g.abs().max().item() == 0 for g in grads)
    g = torch.randn(1, 40, 128, generator=torch.Generator().manual_seed(2)).to(DEV)
    g[0, 3, 5] = float("inf")
    _, dx, grads = _run(att, ffn, x, None, g)
    assert not torch.isfinite(dx).all() or not all(torch.isfinite(t).all() for t in grads)


def test_boundary_tensors_stay_fp32_and_match_the_tf32_engine(stb):
    """Same fp32 tensors in, fp32 tensors out: the two engines agree with each other to the fp32 tolerance on the headline
    width, with input rows whose scales span 1e-2 .. 1 (forward operands are plain fp16: activations are expected inside
    fp16's normal range, as LayerNorm outputs and embeddings are — DESIGN.md §2).  Attention forward + backward and the
    feed-forward forward; the feed-forward backward is only comparable for one ReLU gate pattern and is checked against
    the oracle with the gates pinned (test_gpu_parity.py, both engines)."""
    from speech_tranformer_pytorch_b200 import functional as F
    att, ffn = _layer(stb, seed=8)
    gen = torch.Generator().manual_seed(6)
    x = torch.randn(2, 200, 512, generator=gen)
    x = (x * torch.logspace(-2, 0, 200).view(1, 200, 1)).to(DEV)
    g = torch.randn(2, 200, 512, generator=gen).to(DEV)

    def run():
        for p in att.parameters():
            p.grad = None
        cx = x.clone().requires_grad_()
        a = att(cx, cx, cx)[0]
        a.backward(g)
        with torch.no_grad():
            y = ffn(a.detach())
        return a.detach(), y, cx.grad, [p.grad.clone() for p in att.parameters()]

    a16, y16, dx16, gr16 = run()
    F.set_fp32_engine("tf32")
    a32, y32, dx32, gr32 = run()
    assert all(t.dtype == torch.float32 for t in [a16, y16, dx16] + gr16)
    # two reduced-precision results against each other: twice the bound each has against the fp64 oracle
    assert relerr(a16, a32) < 2 * TOL and relerr(y16, y32) < 2 * TOL and relerr(dx16, dx32) < 2 * TOL
    scale = max(t.abs().max().item() for t in gr32)
    for a, b in zip(gr16, gr32):
        assert (a - b).abs().max().item() / scale < 2 * TOL


def test_head_sizes_other_than_64_use_tf32_operands(stb):
    """Attention with d_k != 64 has no 16-bit kernels: the operator takes the TF32 path whatever the engine says."""
    from speech_tranformer_pytorch_b200 import functional as F
    att, _ = _layer(stb, d=128, H=4, dff=256)          # d_k = 32
    x = torch.randn(2, 70, 128, generator=torch.Generator().manual_seed(3)).to(DEV)
    y16 = att(x, x, x)[0]
    F.set_fp32_engine("tf32")
    assert torch.equal(y16, att(x, x, x)[0])


def test_trainer_keeps_fp16_twins_for_an_fp32_model(stb):
    """DataParallelTrainer(compute_dtype=float32) keeps fp16 operand twins when the composite operators run the fp16-operand
    engine (every head 64 wide), TF32 twins otherwise; either way a training step moves the loss the same way."""
    from speech_tranformer_pytorch_b200 import functional as F
    from speech_tranformer_pytorch_b200 import parallel as spar

    class Net(torch.nn.Module):
        def __init__(self, H):
            super().__init__()
            self.att = stb.MultiHeadAttention(H, 128, 128 // H, 128 // H, dropout=0.0)
            self.ffn = stb.PositionwiseFeedForward(128, 256, dropout=0.0)

        def forward(self, x):
            return self.ffn(self.att(x, x, x)[0])

    torch.manual_seed(0)
    x = torch.randn(2, 50, 128, device=DEV)
    for H, want in ((2, torch.float16), (4, torch.float32)):
        net = Net(H).to(DEV)
        tr = spar.DataParallelTrainer(net, d_model=128, n_warmup_steps=2)
        assert tr.fp.flat_tf32.dtype == want, (H, tr.fp.flat_tf32.dtype)
        losses = []
        for _ in range(4):
            tr.zero_grad()
            loss = (net(x) - 0.5).pow(2).mean()
            loss.backward()
            tr.step()
            losses.append(float(loss))
        assert losses[-1] < losses[0], losses
    F.set_fp32_engine("tf32")
    assert spar.DataParallelTrainer(Net(2).to(DEV), d_model=128).fp.flat_tf32.dtype == torch.float32


def test_c_abi_rejects_the_mixed_code_where_it_has_no_meaning(stb):
    """ST_DTYPE_F32_H16 is a mode of the composite operators only: the primitive entry points reject it."""
    import ctypes as C
    lib = stb._lib.load()
    a = torch.zeros(64, 64, device=DEV)
    rc = lib.st_cast(a.data_ptr(), stb._lib.DTYPE_F32_H16, 64, a.data_ptr(), stb._lib.DTYPE_F32, 64, 64, 64, C.c_float(1.0), None)
    assert rc != 0
    rc = lib.st_gemm_dt(stb._lib.DTYPE_F32_H16, 0, a.data_ptr(), 64, a.data_ptr(), 64, a.data_ptr(), 64, 0, 64, 64, 64, None, None)
    assert rc != 0


def test_chained_operators_skip_conversion_and_amax_passes(stb):
    """An operator's fp32 output carries the fp16 copy its LayerNorm kernel wrote and its fp32 input gradient carries the
    max|.| its last GEMM epilogue measured (functional._H16_ATTR / _AMAX_ATTR): the next operator skips its conversion /
    amax pass.  Same bits as the unchained sequence (a `* 1.0` between the operators drops both tags), fewer launches."""
    att, ffn = _layer(stb, seed=21)
    att2, _ = _layer(stb, seed=22)
    lib = stb._lib.load()
    gen = torch.Generator().manual_seed(23)
    x = torch.randn(2, 130, 512, generator=gen).to(DEV)
    g = torch.randn(2, 130, 512, generator=gen).to(DEV)
    params = list(att.parameters()) + list(ffn.parameters()) + list(att2.parameters())

    def run(chained):
        for p in params:
            p.grad = None
        cx = x.clone().requires_grad_()
        n0 = lib.st_launch_count()
        a = att(cx, cx, cx)[0]
        f = ffn(a if chained else a * 1.0)
        f = f if chained else f * 1.0
        y = att2(f, f, f)[0]
        y.backward(g)
        return y.detach(), cx.grad, [p.grad.clone() for p in params], lib.st_launch_count() - n0

    y1, dx1, gr1, n1 = run(True)
    y0, dx0, gr0, n0 = run(False)
    assert torch.equal(y1, y0) and torch.equal(dx1, dx0)
    for a, b in zip(gr1, gr0):
        assert relerr(a, b) < 2e-4          # split-K / column-sum atomics: fp32 summation order
    # per hand-over: one conversion launch forward, one amax launch backward (two hand-overs here)
    assert n0 - n1 == 4, (n0, n1)


def test_stale_tags_are_not_used(stb):
    """A tag is tied to the tensor's version counter: after an in-place edit of an operator's output the next operator
    converts the edited values itself."""
    from speech_tranformer_pytorch_b200 import functional as F
    att, ffn = _layer(stb, seed=31)
    x = torch.randn(1, 64, 512, generator=torch.Generator().manual_seed(32)).to(DEV)
    with torch.no_grad():
        a = att(x, x, x)[0]
        assert F._tagged(a, F._H16_ATTR) is not None
        want = ffn(a.clone() * 3.0)
        a.mul_(3.0)
        assert F._tagged(a, F._H16_ATTR) is None
        assert torch.equal(ffn(a), want)
        # an input tensor is converted once and the copy reused while it is unchanged
        assert F._tagged(x, F._H16_ATTR) is not None
        n0 = stb._lib.load().st_launch_count()
        att(x, x, x)
        n1 = stb._lib.load().st_launch_count()
        x.add_(1.0)
        att(x, x, x)
        assert stb._lib.load().st_launch_count() - n1 == (n1 - n0) + 1
