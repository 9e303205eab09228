"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (imported from
/root/reference, read-only) on seeded inputs.  Run in the build container only — the reference does
not exist on the GPU box; the fixtures are what travels.

    python oracle/make_golden.py

The reference ships no golden vectors of its own, so these fixtures are what pins
oracle/st_oracle.py (tests/test_oracle_golden.py) and, through it, the CUDA path.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("ST_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# transformer/Utils.py imports packages that are not installed and not used by the hot path
for name in ("editdistance", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, REF)

from transformer.Attention import MultiHeadAttention, ScaledDotProductAttention  # noqa: E402
from transformer.Layers import DecoderLayer, EncoderLayer  # noqa: E402
from transformer.Loss import CrossEntropyLoss, LabelSmoothingLoss  # noqa: E402
from transformer.SubLayers import PositionwiseFeedForward  # noqa: E402
from transformer.Utils import feature_info_mask, padding_info_mask  # noqa: E402


def npy(t):
    return t.detach().cpu().numpy()


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print(f"wrote {name}.npz ({len(arrays)} arrays)")


def params_of(mod, prefix="p."):
    return {prefix + k: npy(v) for k, v in mod.state_dict().items()}


def grads_of(mod, prefix="g."):
    return {prefix + k: npy(p.grad) for k, p in mod.named_parameters()}


def randomize(mod, gen):
    """Non-trivial biases / LayerNorm affine so every parameter gradient is exercised."""
    with torch.no_grad():
        for name, p in mod.named_parameters():
            if p.dim() >= 2:
                torch.nn.init.xavier_normal_(p, generator=gen)
            elif "layernorm.weight" in name:
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=gen))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=gen))


def pad_mask(q_lens, k_lens):
    return padding_info_mask(torch.tensor(q_lens), torch.tensor(k_lens)).bool()   # compat: bool masks


def mha_case(name, n_head, d_model, B, Lq, Lk, mask, seed, cross=False):
    gen = torch.Generator().manual_seed(seed)
    m = MultiHeadAttention(n_head, d_model, d_model // n_head, d_model // n_head, dropout=0.1).eval()
    randomize(m, gen)
    q = torch.randn(B, Lq, d_model, generator=gen, requires_grad=True)
    kv = torch.randn(B, Lk, d_model, generator=gen, requires_grad=True) if cross else q
    out, attn = m(q, kv, kv, mask)
    g = torch.randn(out.shape, generator=gen)
    out.backward(g)
    arrays = dict(q=npy(q), kv=npy(kv), g=npy(g), out=npy(out), attn=npy(attn), dq=npy(q.grad),
                  n_head=np.int64(n_head), cross=np.int64(cross), **params_of(m), **grads_of(m))
    if cross:
        arrays["dkv"] = npy(kv.grad)
    if mask is not None:
        arrays["mask"] = npy(mask.contiguous()).astype(np.uint8)
    save(name, **arrays)


def main():
    torch.manual_seed(2018)
    torch.set_num_threads(1)
    # ---- masks, byte for byte (Utils.py:41-70)
    lens = torch.tensor([5, 3])
    save("masks", lens=npy(lens), pad=npy(padding_info_mask(lens, lens).contiguous()),
         sub=npy(feature_info_mask(lens)),
         pad_qk=npy(padding_info_mask(torch.tensor([4, 2, 3]), torch.tensor([7, 5, 2])).contiguous()))

    # ---- MultiHeadAttention (Attention.py:40-96)
    mha_case("mha_self_padmask", 2, 64, 2, 9, 9, pad_mask([9, 6], [9, 6]), seed=1)
    dec_mask = torch.gt(padding_info_mask(torch.tensor([8, 5]), torch.tensor([8, 5])) +
                        feature_info_mask(torch.tensor([8, 5])), 0)
    mha_case("mha_self_causal", 2, 64, 2, 8, 8, dec_mask, seed=2)
    mha_case("mha_self_nomask_h4", 4, 128, 3, 5, 5, None, seed=3)
    mha_case("mha_cross_eqlen", 2, 64, 2, 7, 7, pad_mask([7, 7], [7, 4]), seed=4, cross=True)

    # ---- ScaledDotProductAttention (Attention.py:9-37)
    gen = torch.Generator().manual_seed(5)
    sd = ScaledDotProductAttention(32).eval()
    q, k, v = (torch.randn(2, L, 32, generator=gen, requires_grad=True) for L in (6, 10, 10))
    m = pad_mask([6, 6], [10, 7])
    o, w = sd(q, k, v, m)
    g = torch.randn(o.shape, generator=gen)
    o.backward(g)
    save("sdpa", q=npy(q), k=npy(k), v=npy(v), mask=npy(m.contiguous()).astype(np.uint8), g=npy(g), out=npy(o),
         attn=npy(w), dq=npy(q.grad), dk=npy(k.grad), dv=npy(v.grad))

    # ---- PositionwiseFeedForward (SubLayers.py:9-28)
    gen = torch.Generator().manual_seed(6)
    ff = PositionwiseFeedForward(64, 128, dropout=0.1).eval()
    randomize(ff, gen)
    x = torch.randn(2, 9, 64, generator=gen, requires_grad=True)
    y = ff(x)
    g = torch.randn(y.shape, generator=gen)
    y.backward(g)
    save("ffn", x=npy(x), g=npy(g), out=npy(y), dx=npy(x.grad), **params_of(ff), **grads_of(ff))

    # ---- EncoderLayer / DecoderLayer (Layers.py:8-44) — the reference's own composition
    gen = torch.Generator().manual_seed(7)
    enc = EncoderLayer(64, 128, 2, 32, 32, dropout=0.1).eval()
    randomize(enc, gen)
    x = torch.randn(2, 11, 64, generator=gen, requires_grad=True)
    m = pad_mask([11, 7], [11, 7])
    y, _ = enc(x, m)
    g = torch.randn(y.shape, generator=gen)
    y.backward(g)
    save("encoder_layer", x=npy(x), mask=npy(m.contiguous()).astype(np.uint8), g=npy(g), out=npy(y), dx=npy(x.grad),
         **params_of(enc), **grads_of(enc))

    # DecoderLayer as written only runs when len_q == len_k (residual adds v, Attention.py:94)
    gen = torch.Generator().manual_seed(8)
    dec = DecoderLayer(64, 128, 2, 32, 32, dropout=0.1).eval()
    randomize(dec, gen)
    x = torch.randn(2, 6, 64, generator=gen, requires_grad=True)
    e = torch.randn(2, 6, 64, generator=gen, requires_grad=True)
    sm = torch.gt(padding_info_mask(torch.tensor([6, 4]), torch.tensor([6, 4])) + feature_info_mask(torch.tensor([6, 4])), 0)
    em = pad_mask([6, 4], [6, 5])
    y, _ = dec(x, e, sm, em)
    g = torch.randn(y.shape, generator=gen)
    y.backward(g)
    save("decoder_layer_eqlen", x=npy(x), enc=npy(e), slf_mask=npy(sm).astype(np.uint8),
         enc_mask=npy(em.contiguous()).astype(np.uint8), g=npy(g), out=npy(y), dx=npy(x.grad), denc=npy(e.grad),
         **params_of(dec), **grads_of(dec))

    # ---- LabelSmoothingLoss / CrossEntropyLoss (Loss.py)
    gen = torch.Generator().manual_seed(9)
    N, V = 12, 30
    logits = torch.randn(N, V, generator=gen)
    target = torch.randint(0, V, (N,), generator=gen)
    target[2] = 0
    target[5] = 3
    target[7] = 0
    weight = 0.5 + torch.rand(V, generator=gen)
    arrays = dict(logits=npy(logits), target=npy(target), weight=npy(weight))
    for ii in (-1, 0, 3):
        for sa in (True, False):
            for w, wname in ((weight, "w"), (torch.ones(V), "u")):
                x = logits.clone().requires_grad_()
                crit = LabelSmoothingLoss(0.1, V, weight=w, size_average=sa, ignore_index=ii)
                loss = crit(x, target)
                loss.backward()
                key = f"ii{ii}_sa{int(sa)}_{wname}"
                arrays["loss_" + key] = npy(loss)
                arrays["grad_" + key] = npy(x.grad)
                arrays["onehot_" + key] = npy(crit.one_hot)
    q = torch.softmax(torch.randn(N, V, generator=gen), -1)
    x = logits.clone().requires_grad_()
    loss = CrossEntropyLoss(weight, size_average=True)(x, q)
    loss.backward()
    arrays.update(dense_q=npy(q), dense_loss=npy(loss), dense_grad=npy(x.grad))
    save("lsce", **arrays)


if __name__ == "__main__":
    main()
