"""The CPU oracle (oracle/st_oracle.py) against fixtures produced by the reference modules themselves
(oracle/make_golden.py -> tests/golden).  fp32 restatement of the same ATen ops => agreement ~1e-6."""
import numpy as np
import pytest
import torch

from helpers import golden, relerr, t, torch_params
from oracle import st_oracle as O

TIGHT = 2e-5


def test_masks_byte_identical():
    g = golden("masks")
    lens = t(g["lens"])
    assert np.array_equal(O.padding_info_mask(lens, lens).contiguous().numpy(), g["pad"])
    assert np.array_equal(O.feature_info_mask(lens).numpy(), g["sub"])
    assert np.array_equal(O.padding_info_mask(torch.tensor([4, 2, 3]), torch.tensor([7, 5, 2])).contiguous().numpy(),
                          g["pad_qk"])
    m = O.padding_info_mask(lens, lens)
    assert m.stride(1) == 0, "key-padding mask must stay a stride-0 expanded view like Utils.py:53-54"


@pytest.mark.parametrize("name", ["mha_self_padmask", "mha_self_causal", "mha_self_nomask_h4", "mha_cross_eqlen"])
def test_mha(name):
    g = golden(name)
    P = torch_params(g, requires_grad=True)
    q = t(g["q"]).requires_grad_()
    cross = bool(g["cross"])
    kv = t(g["kv"]).requires_grad_() if cross else q
    mask = t(g["mask"]).bool() if "mask" in g else None
    out, attn = O.multi_head_attention(q, kv, kv, mask, P, int(g["n_head"]))
    out.backward(t(g["g"]))
    assert relerr(out, g["out"]) < TIGHT
    assert relerr(attn, g["attn"]) < TIGHT
    assert relerr(q.grad, g["dq"]) < TIGHT
    if cross:
        assert relerr(kv.grad, g["dkv"]) < TIGHT
    scale = max(np.abs(g["g." + k]).max() for k in P)
    for k, p in P.items():  # d(bias_k) is analytically 0: compare on the common gradient scale
        assert (p.grad - t(g["g." + k])).abs().max().item() / scale < TIGHT, k
    if mask is not None:
        assert torch.equal(attn[mask.unsqueeze(1).expand_as(attn)], torch.zeros_like(attn)[mask.unsqueeze(1).expand_as(attn)])


def test_sdpa():
    g = golden("sdpa")
    q, k, v = (t(g[n]).requires_grad_() for n in "qkv")
    out, w = O.scaled_dot_product_attention(q, k, v, t(g["mask"]).bool(), 32)
    out.backward(t(g["g"]))
    for a, b in ((out, "out"), (w, "attn"), (q.grad, "dq"), (k.grad, "dk"), (v.grad, "dv")):
        assert relerr(a, g[b]) < TIGHT, b


def test_ffn():
    g = golden("ffn")
    P = torch_params(g, requires_grad=True)
    x = t(g["x"]).requires_grad_()
    y = O.positionwise_ffn(x, P)
    y.backward(t(g["g"]))
    assert relerr(y, g["out"]) < TIGHT
    assert relerr(x.grad, g["dx"]) < TIGHT
    for k, p in P.items():
        assert relerr(p.grad, g["g." + k]) < TIGHT, k


def test_encoder_layer():
    g = golden("encoder_layer")
    P = torch_params(g, requires_grad=True)
    x = t(g["x"]).requires_grad_()
    y = O.encoder_layer(x, t(g["mask"]).bool(), P, 2)
    y.backward(t(g["g"]))
    assert relerr(y, g["out"]) < TIGHT
    assert relerr(x.grad, g["dx"]) < TIGHT
    scale = max(np.abs(g["g." + k]).max() for k in P)
    for k, p in P.items():
        assert (p.grad - t(g["g." + k])).abs().max().item() / scale < TIGHT, k


def test_decoder_layer_reference_residual():
    """Layers.py:37-44 as written (cross-attention residual adds the ENCODER output; len_q == len_k)."""
    g = golden("decoder_layer_eqlen")
    P = torch_params(g, requires_grad=True)
    x, e = t(g["x"]).requires_grad_(), t(g["enc"]).requires_grad_()
    sub = lambda pre: {k[len(pre):]: v for k, v in P.items() if k.startswith(pre)}
    a, _ = O.multi_head_attention(x, x, x, t(g["slf_mask"]).bool(), sub("slf_attn."), 2)
    c, _ = O.multi_head_attention(a, e, e, t(g["enc_mask"]).bool(), sub("enc_attn."), 2, residual="v")
    y = O.positionwise_ffn(c, sub("pos_ffn."))
    y.backward(t(g["g"]))
    assert relerr(y, g["out"]) < TIGHT
    assert relerr(x.grad, g["dx"]) < TIGHT
    assert relerr(e.grad, g["denc"]) < TIGHT


@pytest.mark.parametrize("ii", [-1, 0, 3])
@pytest.mark.parametrize("sa", [True, False])
@pytest.mark.parametrize("wname", ["w", "u"])
def test_label_smoothing_loss(ii, sa, wname):
    g = golden("lsce")
    key = f"ii{ii}_sa{int(sa)}_{wname}"
    V = g["logits"].shape[1]
    w = t(g["weight"]) if wname == "w" else torch.ones(V)
    one_hot = O.smoothing_one_hot(0.1, V, ii)
    assert np.array_equal(one_hot.numpy(), g["onehot_" + key])
    x = t(g["logits"]).requires_grad_()
    loss = O.label_smoothing_loss(x, t(g["target"]), one_hot, w, 0.1, ii, sa)
    loss.backward()
    assert relerr(loss, g["loss_" + key]) < TIGHT
    assert relerr(x.grad, g["grad_" + key]) < TIGHT


def test_soft_cross_entropy_dense():
    g = golden("lsce")
    x = t(g["logits"]).requires_grad_()
    loss = O.soft_cross_entropy(x, t(g["dense_q"]), t(g["weight"]), True)
    loss.backward()
    assert relerr(loss, g["dense_loss"]) < TIGHT
    assert relerr(x.grad, g["dense_grad"]) < TIGHT


def test_closed_form_gradient():
    """SURVEY §8a closed form: dL/dx = (1/Z) [softmax(x) * sum_c w q - w q] — what the CUDA kernel computes."""
    g = golden("lsce")
    x, tg, w = t(g["logits"]).double(), t(g["target"]), t(g["weight"]).double()
    one_hot = O.smoothing_one_hot(0.1, x.shape[1], 0, torch.float64)
    q = one_hot.repeat(x.shape[0], 1)
    q.scatter_(1, tg.unsqueeze(1), 0.9)
    q.masked_fill_((tg == 0).unsqueeze(1), 0)
    grad = (torch.softmax(x, -1) * (w * q).sum(-1, keepdim=True) - w * q) / x.shape[0]
    assert relerr(grad, g["grad_ii0_sa1_w"]) < TIGHT


def test_synthetic_batch_layout():
    x, tgt, il, tl, gt = O.synthetic_batch(4, 50, 12, 80, 100, fixed_len=False, t_min=10, l_min=3)
    assert x.shape == (4, 50, 80) and tgt.shape == (4, 12) and gt.shape == (4, 12)
    for b in range(4):
        assert torch.all(x[b, int(il[b]):] == 0)
        n = int(tl[b])
        assert tgt[b, 0] == O.BOS and torch.all(tgt[b, n:] == O.PAD) and gt[b, n - 1] == O.BOS
        assert torch.equal(tgt[b, 1:n], gt[b, :n - 1])


# ------------------------------------------------------------------------------------------------ whole model
def test_model_port_matches_reference_model():
    """oracle/model_port.py (front-end, embedding, layer stacks, vocabulary projection, loss) against the
    reference's own Transformer run by oracle/make_golden_model.py (eval mode, ragged batch): logits, loss and
    every parameter gradient."""
    from oracle import model_port
    g = golden("transformer_small")
    cfg = dict(d_model=64, n_heads=2, num_enc_layer=2, num_dec_layer=2, vocab_size=31)
    P = {k[2:]: t(v).clone().requires_grad_() for k, v in g.items() if k.startswith("p.") and not k.endswith(".pe")}
    inputs, targets = t(g["inputs"]), t(g["targets"])
    in_len, tgt_len, truth = t(g["in_len"]), t(g["tgt_len"]), t(g["truth"])
    logits = model_port.forward(P, cfg, inputs, in_len, targets, tgt_len)
    V = 31
    loss = O.label_smoothing_loss(logits.reshape(-1, V), truth.reshape(-1), O.smoothing_one_hot(0.1, V, 0), torch.ones(V),
                                  0.1, 0, True)
    loss.backward()
    assert relerr(logits, g["logits"]) < 2e-5
    assert abs(float(loss) - float(g["loss"])) < 2e-5 * abs(float(g["loss"]))
    scale = max(np.abs(v).max() for k, v in g.items() if k.startswith("g."))
    for k, v in g.items():
        if k.startswith("g."):
            got = P[k[2:]].grad
            assert got is not None, k
            assert (got.numpy() - v).__abs__().max() / scale < 2e-5, k


def test_decode_port_beam1_is_greedy_and_scores_are_sorted():
    """oracle/decode_port.py self-consistency on the reference-pinned model port: width 1 equals greedy arg-max
    decoding with the summed log-probability as its score; wider beams return scores in descending order that are
    at least as good as greedy."""
    from oracle import decode_port, model_port
    g = golden("transformer_small")
    cfg = dict(d_model=64, n_heads=2, num_enc_layer=2, num_dec_layer=2, vocab_size=31)
    P = {k[2:]: t(v).double() for k, v in g.items() if k.startswith("p.") and not k.endswith(".pe")}
    P["tgt_word_proj.weight"] = P["tgt_word_proj.weight"] * 12
    inputs, in_len = t(g["inputs"]).double(), t(g["in_len"])
    hyps, scores = decode_port.beam_search(P, cfg, inputs, in_len, beam=1, max_len=6)
    for b in range(inputs.size(0)):
        prefix, total = [decode_port.BOS], 0.0
        for _ in range(6):
            lp = torch.log_softmax(decode_port.step_logits(P, cfg, inputs[b:b + 1, :int(in_len[b])], in_len[b:b + 1], torch.tensor([prefix])), -1)[0]
            y = int(lp.argmax())
            total += float(lp[y])
            prefix.append(y)
            if y == decode_port.EOS:
                break
        assert hyps[b][0] == prefix[1:]
        assert abs(float(scores[b, 0]) - total) < 1e-9
    wide_h, wide_s = decode_port.beam_search(P, cfg, inputs, in_len, beam=4, max_len=6, n_best=4)
    assert bool((wide_s[:, :-1] >= wide_s[:, 1:]).all())
    assert bool((wide_s[:, 0] >= scores[:, 0] - 1e-9).all())


# ------------------------------------------------------------------------------------------------ beam search bookkeeping
def test_beam_port_matches_reference_beam_class():
    """oracle/decode_port.beam_advance / hypothesis against what the reference's own Beam class (Beam.py:43-74,99-116) kept for
    seeded log-probability tables (oracle/make_golden_beam.py): scores to 1e-6, back-pointers / symbols / hypotheses exactly.
    This pins the oracle that the GPU beam search (decode.py, st_beam_step) is compared with."""
    from oracle import decode_port
    g = golden("beam_advance")
    assert tuple(g["constants"]) == (decode_port.PAD, decode_port.UNK, decode_port.BOS, decode_port.EOS)
    for ci, (beam, V, steps, seed) in enumerate(g["cases"].tolist()):
        lk, want_s, want_p, want_y = (g[f"c{ci}.{k}"] for k in ("word_lk", "scores", "prev_k", "next_y"))
        scores = torch.zeros(1, beam)
        prev_ks, next_ys = [], []
        for t in range(lk.shape[0]):
            scores, prev_k, y = decode_port.beam_advance(scores, torch.from_numpy(lk[t]).unsqueeze(0), first=(t == 0))
            assert np.array_equal(prev_k[0].numpy(), want_p[t]) and np.array_equal(y[0].numpy(), want_y[t]), (ci, t)
            assert np.abs(scores[0].numpy() - want_s[t]).max() < 1e-6
            prev_ks.append(prev_k[0])
            next_ys.append(y[0])
            assert bool(y[0, 0] == decode_port.EOS) == bool(g[f"c{ci}.done"][t])          # Beam.py:70-72
        for k in range(beam):
            assert decode_port.hypothesis(prev_ks, next_ys, k) == g[f"c{ci}.hyps"][k].tolist()
