"""GPU parity of the CTC head against oracle.st_oracle.ctc_nll (PyTorch's CPU ctc_loss in float64 —
the reference's train_attn_and_ctc.py is empty, see the oracle's header)."""
import pytest
import torch

from helpers import relerr
from oracle import st_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def stb():
    import speech_tranformer_pytorch_b200 as m
    m.build()
    m._lib.check(m._lib.load().st_device_check(0))
    return m


def _case(B, T, V, L_max, seed, blank=0, repeats=True):
    g = torch.Generator().manual_seed(seed)
    logits = 2.0 * torch.randn(B, T, V, generator=g)
    in_len = torch.randint(max(T // 2, 2 * L_max + 1), T + 1, (B,), generator=g)
    in_len[0] = T
    tgt_len = torch.randint(1, L_max + 1, (B,), generator=g)
    tgt_len[0] = L_max
    labels = [c for c in range(V) if c != blank]
    targets = torch.tensor(labels)[torch.randint(0, len(labels), (B, L_max), generator=g)]
    if repeats and L_max >= 3:
        targets[:, 1] = targets[:, 0]              # repeated labels need the mandatory blank between them
    return logits, targets, in_len, tgt_len


@pytest.mark.parametrize("B,T,V,L_max,blank", [(3, 30, 11, 6, 0), (2, 9, 5, 4, 4), (4, 300, 4337, 50, 0), (5, 64, 31, 1, 0)])
def test_ctc_matches_oracle(stb, B, T, V, L_max, blank):
    F = stb.functional
    logits, targets, in_len, tgt_len = _case(B, T, V, L_max, seed=B * 100 + T, blank=blank)
    w = torch.linspace(0.5, 1.5, B)                                   # a non-trivial upstream gradient per utterance
    cl = logits.to(DEV).requires_grad_()
    nll = F.ctc_loss(cl, targets.to(DEV), in_len.to(DEV), tgt_len.to(DEV), blank=blank, reduction="none")
    (nll * w.to(DEV)).sum().backward()
    rl = logits.double().requires_grad_()
    ref = O.ctc_nll(rl, targets, in_len, tgt_len, blank)
    (ref * w.double()).sum().backward()
    assert relerr(nll, ref) < 1e-5
    assert relerr(cl.grad, rl.grad) < 1e-4
    for b in range(B):                                                # frames beyond the utterance: exactly zero
        assert torch.count_nonzero(cl.grad[b, int(in_len[b]):]).item() == 0


def test_ctc_reductions_and_module(stb):
    logits, targets, in_len, tgt_len = _case(4, 40, 13, 5, seed=7)
    ref = O.ctc_nll(logits.double(), targets, in_len, tgt_len)
    args = (logits.to(DEV), targets.to(DEV), in_len.to(DEV), tgt_len.to(DEV))
    assert relerr(stb.CTCLoss(reduction="sum")(*args), ref.sum()) < 1e-5
    assert relerr(stb.CTCLoss(reduction="mean")(*args), (ref / tgt_len.double()).mean()) < 1e-5
    assert relerr(stb.CTCLoss(reduction="none")(*args), ref) < 1e-5


def test_ctc_infeasible_and_empty_targets(stb):
    F = stb.functional
    B, T, V = 3, 6, 7
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(B, T, V, generator=g)
    targets = torch.tensor([[1, 1, 2, 3], [2, 3, 0, 0], [0, 0, 0, 0]])
    in_len = torch.tensor([4, 6, 5])           # utterance 0: 4 labels incl. a repeat need >= 5 frames -> no alignment
    tgt_len = torch.tensor([4, 2, 0])          # utterance 2: empty target -> all-blank path
    cl = logits.to(DEV).requires_grad_()
    nll = F.ctc_loss(cl, targets.to(DEV), in_len.to(DEV), tgt_len.to(DEV), reduction="none")
    ref = O.ctc_nll(logits.double(), targets, in_len, tgt_len)
    assert torch.isinf(nll[0]) and torch.isinf(ref[0])
    assert relerr(nll[1:], ref[1:]) < 1e-5
    nll[1:].sum().backward()
    assert torch.count_nonzero(cl.grad[0]).item() == 0 and torch.isfinite(cl.grad).all()


def test_joint_ctc_attention_step(stb):
    """configs[3] wiring: encoder output -> CTC projection (row-padded logits read in place) + decoder logits -> label-smoothed CE."""
    from test_gpu_model import _small_model
    from helpers import golden, t
    g = golden("transformer_small")
    V = 31
    net = _small_model(stb, g).train()
    ctc_proj = torch.nn.Linear(64, V).to(DEV)
    batch = [t(g[k], DEV) for k in ("inputs", "in_len", "targets", "tgt_len")]
    truth = t(g["truth"], DEV)
    att = stb.LabelSmoothingLoss(0.1, V, weight=torch.ones(V, device=DEV), size_average=True, ignore_index=0).to(DEV)
    crit = stb.JointCTCAttentionLoss(att, ctc_weight=0.3, blank=0)
    enc, _ = net.encoder(batch[0], batch[1])
    dec, _, _ = net.decoder(batch[2], batch[3], batch[1], enc)
    dec_logits = stb.functional.linear(dec, net.tgt_word_proj.weight)
    ctc_logits = stb.functional.linear(enc, ctc_proj.weight, ctc_proj.bias)
    assert not ctc_logits.is_contiguous()                               # V = 31: padded rows, consumed without a copy
    labels = torch.where(truth > 3, truth, torch.full_like(truth, 4))   # any non-blank labels of the right lengths
    loss = crit(dec_logits.view(-1, V), truth.view(-1), ctc_logits, labels, batch[1], batch[3] - 1)
    loss.backward()
    assert torch.isfinite(loss)
    for p in list(net.encoder.parameters()) + list(ctc_proj.parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all()
