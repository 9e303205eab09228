"""CPU ORACLE — test infrastructure only.

A from-scratch restatement, op by op, of the reference's algorithm for the hot path
(ZhengkunTian/Speech-Tranformer-Pytorch: transformer/Attention.py, SubLayers.py, Loss.py, Layers.py,
and the mask builders of Utils.py).  The reference's arithmetic lives in PyTorch ATen (a third-party
dependency the reference does not pin), so this oracle is written with the same CPU tensor ops in
plain functional form: explicit parameters, no nn.Module state, any float dtype (float64 for tight
checks).  Gradients come from torch.autograd over these functions.

PINNING: the reference ships no golden vectors or known-answer tests for this path (SURVEY.md §4,
§8c).  The oracle is therefore pinned against outputs of the reference modules themselves, executed
in the build container by oracle/make_golden.py and committed as tests/golden/*.npz; the
`-m "not gpu"` tests check this file against those fixtures.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.  The
product (speech-tranformer-pytorch_b200/) never does.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
PAD, UNK, BOS, EOS = 0, 1, 2, 3  # transformer/Constants.py:1-4


# ---------------------------------------------------------------------------------------------
# masks — transformer/Utils.py:41-70 (nonzero / True = masked)
# ---------------------------------------------------------------------------------------------
def padding_info_mask(seq_q_length: Tensor, seq_k_length: Tensor) -> Tensor:
    """Utils.py:41-57. (B, max_q, max_k) key-padding mask, an expanded (stride-0 over q) view."""
    assert seq_q_length.dim() == 1 and seq_k_length.dim() == 1
    batch = seq_k_length.size(0)
    len_q = int(seq_q_length.max())
    len_k = int(seq_k_length.max())
    rows = np.zeros((batch, len_k), dtype=np.uint8)
    for i in range(batch):
        rows[i, int(seq_k_length[i]):] = 1
    return torch.from_numpy(rows).unsqueeze(1).expand(batch, len_q, len_k)


def feature_info_mask(seq_length: Tensor) -> Tensor:
    """Utils.py:60-70. (B, L, L) strictly-upper-triangular 'subsequent' mask."""
    assert seq_length.dim() == 1
    batch = seq_length.size(0)
    max_len = int(seq_length.max())
    return torch.from_numpy(np.triu(np.ones((batch, max_len, max_len)), k=1).astype("uint8"))


def decoder_self_mask(target_lengths: Tensor) -> Tensor:
    """Models.py:89-94 with lengths in place of the (broken) token ids: pad mask OR subsequent mask."""
    return torch.gt(padding_info_mask(target_lengths, target_lengths) + feature_info_mask(target_lengths), 0)


# ---------------------------------------------------------------------------------------------
# (a-4) residual + LayerNorm — Attention.py:62,94 ; SubLayers.py:18,27
# ---------------------------------------------------------------------------------------------
def add_layer_norm(a: Tensor, b: Optional[Tensor], gamma: Tensor, beta: Tensor, eps: float = 1e-6) -> Tensor:
    z = a if b is None else a + b
    mean = z.mean(dim=-1, keepdim=True)
    var = ((z - mean) ** 2).mean(dim=-1, keepdim=True)  # biased, as nn.LayerNorm
    return (z - mean) / torch.sqrt(var + eps) * gamma + beta


# ---------------------------------------------------------------------------------------------
# (a-2) ScaledDotProductAttention.forward — Attention.py:17-37
# ---------------------------------------------------------------------------------------------
def scaled_dot_product_attention(q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor], d_k: int
                                 ) -> Tuple[Tensor, Tensor]:
    attn = torch.bmm(q, k.transpose(1, 2)) / math.sqrt(d_k)           # :27
    if mask is not None:
        assert mask.size() == attn.size()                             # :30
        attn = attn.masked_fill(mask.bool(), -float("inf"))           # :31
    weights = torch.softmax(attn, dim=-1)                             # :33  (dropout p=0, :34)
    return torch.bmm(weights, v), weights                             # :35


# ---------------------------------------------------------------------------------------------
# (a-1) MultiHeadAttention.forward — Attention.py:64-96
# params: linear_q|linear_k|linear_v|output_linear .weight/.bias, layernorm.weight/.bias
# ---------------------------------------------------------------------------------------------
def multi_head_attention(q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor], params: Dict[str, Tensor],
                         n_head: int, residual: str = "v", eps: float = 1e-6) -> Tuple[Tensor, Tensor]:
    B, d_model = q.size(0), q.size(-1)
    d_k = d_model // n_head

    def lin(x, name):
        return x @ params[name + ".weight"].t() + params[name + ".bias"]

    def split(x):  # :68-69
        return x.view(B, -1, n_head, d_k).transpose(1, 2)

    query, key, value = split(lin(q, "linear_q")), split(lin(k, "linear_k")), split(lin(v, "linear_v"))  # :74-80
    scores = torch.matmul(query, key.transpose(2, 3)) / math.sqrt(d_k)                                   # :82
    if mask is not None:
        scores = scores.masked_fill(mask.bool().unsqueeze(1), -float("inf"))                             # :84-87
    attns = torch.softmax(scores, dim=-1)                                                                # :89 (p=0)
    context = torch.matmul(attns, value).transpose(1, 2).contiguous().view(B, -1, n_head * d_k)          # :90
    output = lin(context, "output_linear")                                                               # :92
    res = v if residual == "v" else q                                                                    # :94 adds v
    out = add_layer_norm(output, res, params["layernorm.weight"], params["layernorm.bias"], eps)
    return out, attns


# ---------------------------------------------------------------------------------------------
# (a-3) PositionwiseFeedForward.forward — SubLayers.py:24-28
# params: fc1|fc2 .weight/.bias, layernorm.weight/.bias
# ---------------------------------------------------------------------------------------------
def ffn_preactivation(x: Tensor, params: Dict[str, Tensor]) -> Tensor:
    return x @ params["fc1.weight"].t() + params["fc1.bias"]


def positionwise_ffn(x: Tensor, params: Dict[str, Tensor], eps: float = 1e-6, gate: Optional[Tensor] = None) -> Tensor:
    """`gate` (0/1, shape of the hidden activation) overrides the ReLU's own active set.  A test compares
    gradients for a FIXED gate pattern, because a reduced-precision forward may legitimately put a
    pre-activation that is within rounding error of 0 on the other side of the kink."""
    pre = ffn_preactivation(x, params)
    h = torch.relu(pre) if gate is None else pre * gate.to(pre.dtype)     # :25
    y = h @ params["fc2.weight"].t() + params["fc2.bias"]                  # :26
    return add_layer_norm(x, y, params["layernorm.weight"], params["layernorm.bias"], eps)  # :27


# ---------------------------------------------------------------------------------------------
# Layers.py:18-22 / :37-44
# ---------------------------------------------------------------------------------------------
def _sub(params: Dict[str, Tensor], prefix: str) -> Dict[str, Tensor]:
    n = len(prefix)
    return {k[n:]: v for k, v in params.items() if k.startswith(prefix)}


def encoder_layer(x: Tensor, mask: Optional[Tensor], params: Dict[str, Tensor], n_head: int,
                  ffn_gate: Optional[Tensor] = None) -> Tensor:
    a, _ = multi_head_attention(x, x, x, mask, _sub(params, "slf_attn."), n_head)
    return positionwise_ffn(a, _sub(params, "pos_ffn."), gate=ffn_gate)


def decoder_layer(x: Tensor, enc: Tensor, slf_mask: Optional[Tensor], enc_mask: Optional[Tensor],
                  params: Dict[str, Tensor], n_head: int, residual: str = "q", ffn_gate: Optional[Tensor] = None) -> Tensor:
    a, _ = multi_head_attention(x, x, x, slf_mask, _sub(params, "slf_attn."), n_head)
    c, _ = multi_head_attention(a, enc, enc, enc_mask, _sub(params, "enc_attn."), n_head, residual=residual)
    return positionwise_ffn(c, _sub(params, "pos_ffn."), gate=ffn_gate)


# ---------------------------------------------------------------------------------------------
# (a-6) CrossEntropyLoss.forward (soft target) — Loss.py:50-73
# ---------------------------------------------------------------------------------------------
def soft_cross_entropy(inputs: Tensor, target: Tensor, weight: Tensor, size_average: bool = True) -> Tensor:
    assert inputs.dim() == 2 and target.dim() == 2                    # :51-52
    logp = torch.log_softmax(inputs, dim=-1)                          # :57
    w = weight.unsqueeze(0).expand_as(inputs)                         # :59
    tmp = -(logp * target)                                            # :61-62
    weighted = w * tmp                                                # :64-65
    loss = weighted.sum()
    return loss / inputs.size(0) if size_average else loss            # :67-71


# ---------------------------------------------------------------------------------------------
# (a-5) LabelSmoothingLoss — Loss.py:13-39
# ---------------------------------------------------------------------------------------------
def smoothing_one_hot(label_smoothing: float, vocab_size: int, ignore_index: int, dtype=torch.float32) -> Tensor:
    """Loss.py:18-22: the (1, V) `one_hot` buffer. Column ignore_index is zeroed only when it is 0."""
    assert 0.0 <= label_smoothing <= 1.0                              # :14
    one_hot = torch.full((vocab_size,), label_smoothing / (vocab_size - 1), dtype=dtype)
    if not ignore_index:                                              # :20 (sic)
        one_hot[ignore_index] = 0
    return one_hot.unsqueeze(0)


def label_smoothing_loss(output: Tensor, target: Tensor, one_hot: Tensor, weight: Tensor, label_smoothing: float,
                         ignore_index: int = -1, size_average: bool = True) -> Tensor:
    confidence = 1.0 - label_smoothing                                # :24
    model_prob = one_hot.to(output.dtype).repeat(target.size(0), 1)   # :33
    model_prob.scatter_(1, target.unsqueeze(1), confidence)           # :34
    if ignore_index >= 0:                                             # :35-37
        model_prob.masked_fill_((target == ignore_index).unsqueeze(1), 0)
    return soft_cross_entropy(output, model_prob, weight, size_average)


# ---------------------------------------------------------------------------------------------
# synthetic data in the layout of Dataset.__getitem__ (Dataset.py:34-51) — SURVEY.md §8(d)
# ---------------------------------------------------------------------------------------------
def synthetic_batch(batch: int, t_max: int, l_max: int, feat: int, vocab: int, seed: int = 2018,
                    fixed_len: bool = True, t_min: int = 0, l_min: int = 10):
    g = torch.Generator().manual_seed(seed)
    in_len = torch.full((batch,), t_max, dtype=torch.int64) if fixed_len else \
        torch.randint(max(t_min, 1), t_max + 1, (batch,), generator=g)
    if not fixed_len:
        in_len[0] = t_max
    tgt_len = torch.randint(l_min, l_max + 1, (batch,), generator=g)
    tgt_len[0] = l_max
    inputs = torch.randn(batch, t_max, feat, generator=g)
    targets = torch.zeros(batch, l_max, dtype=torch.int64)
    truth = torch.zeros(batch, l_max, dtype=torch.int64)
    for b in range(batch):
        inputs[b, int(in_len[b]):] = 0
        n = int(tgt_len[b]) - 1
        labels = torch.randint(4, vocab, (n,), generator=g)
        targets[b, 0] = BOS
        targets[b, 1:n + 1] = labels            # [BOS] + labels
        truth[b, :n] = labels
        truth[b, n] = BOS                       # labels + [BOS] (sic, Dataset.py:36-37)
    return inputs, targets, in_len, tgt_len, truth


# ---------------------------------------------------------------------------------------------
# CTC.  The reference's train_attn_and_ctc.py is an EMPTY file: there is no reference code for
# this row.  The algorithm lives in a third-party dependency the reference would have called — PyTorch ATen's
# ctc_loss (torch 2.11.0 in this image; Graves et al. 2006 forward-backward in the log domain).  The oracle is that
# implementation run on CPU in float64; parity of the CUDA kernels is anchored on it (tests/test_gpu_ctc.py).
# ---------------------------------------------------------------------------------------------
def ctc_nll(logits: Tensor, targets: Tensor, input_lengths: Tensor, target_lengths: Tensor, blank: int = 0) -> Tensor:
    """Per-utterance -log p(targets | logits); logits (B, T, V) unnormalised, targets (B, L_max) padded."""
    logp = torch.log_softmax(logits, dim=-1).transpose(0, 1)       # (T, B, V) as F.ctc_loss expects
    return torch.nn.functional.ctc_loss(logp, targets, input_lengths, target_lengths, blank=blank, reduction="none",
                                        zero_infinity=False)
