"""B200 drop-in for the reference's transformer/Attention.py.

Same classes, constructor signatures, sub-module / parameter names (so `state_dict`s interchange:
linear_q|linear_k|linear_v|output_linear.{weight,bias}, layernorm.{weight,bias}) and forward
contracts as the reference (Attention.py:9-37, :40-96).  The arithmetic runs in libst_b200.so.

Deliberate, switchable deviations:
  * `residual`: 'v' reproduces the reference (`output + v`, Attention.py:94, which only works when
    len_q == len_k); 'q' is what a functioning decoder cross-attention needs.
  * `return_attention`: the reference always returns the (B, h, Lq, Lk) weights; materialising them
    costs 1 GB per layer at the headline shape and every caller in Models.py discards them unless
    `return_attns` is set, so the default here is None.  Set the flag to get them.
"""
import math

import torch
import torch.nn as nn

from .. import functional as F

__all__ = ["ScaledDotProductAttention", "MultiHeadAttention"]


class ScaledDotProductAttention(nn.Module):
    """Reference: Attention.py:9-37.  q/k/v are [batch, time, d_k]; mask must match the scores' shape."""

    def __init__(self, d_k, dropout=0):
        super(ScaledDotProductAttention, self).__init__()
        self.d_k = d_k
        self.scaled = math.sqrt(d_k)
        self.softmax = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout)

    def forward(self, q, k, v, mask=None):
        if mask is not None:
            assert mask.size() == (q.size(0), q.size(1), k.size(1))       # Attention.py:30
        if q.size(-1) != self.d_k or q.size(-1) not in (32, 64, 128):
            raise RuntimeError("ScaledDotProductAttention on B200 supports d_k in {32, 64, 128} equal to the last dim")
        p = self.dropout.p if self.training else 0.0
        out, attn = F.attention_core(q, k, v, mask, n_head=1, dropout_p=p, seed=F.next_seed() if p > 0 else 0,
                                     need_attn=True)
        return out, attn.squeeze(1)


class MultiHeadAttention(nn.Module):
    """Reference: Attention.py:40-96."""

    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1, residual="v", return_attention=False):
        super(MultiHeadAttention, self).__init__()

        assert d_model % n_head == 0                  # Attention.py:45-47
        assert d_v == int(d_model / n_head)
        assert d_k == int(d_model / n_head)

        self.d_model = d_model
        self.n_head = n_head
        self.d_k = d_k
        self.d_v = d_v
        self.scaled = math.sqrt(d_k)
        self.residual = residual
        self.return_attention = return_attention

        self.linear_q = nn.Linear(d_model, n_head * d_k)
        self.linear_k = nn.Linear(d_model, n_head * d_k)
        self.linear_v = nn.Linear(d_model, n_head * d_v)

        self.softmax = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout)
        self.output_linear = nn.Linear(d_model, d_model)
        self.layernorm = nn.LayerNorm(d_model, eps=1e-6)

    def forward(self, q, k, v, mask=None):
        p = self.dropout.p if self.training else 0.0
        out, attns = F.multi_head_attention(
            q, k, v, mask,
            self.linear_q.weight, self.linear_q.bias, self.linear_k.weight, self.linear_k.bias,
            self.linear_v.weight, self.linear_v.bias, self.output_linear.weight, self.output_linear.bias,
            self.layernorm.weight, self.layernorm.bias,
            n_head=self.n_head, residual=self.residual, eps=self.layernorm.eps, dropout_p=p,
            seed=F.next_seed() if p > 0 else 0, need_attn=self.return_attention, round_out=F.ROUND_OUT)
        return out, attns
