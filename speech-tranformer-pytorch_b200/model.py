"""Model assembly around the B200 hot-path modules — the callers either side of the path.

Mirrors transformer/Layers.py (EncoderLayer :8-22, DecoderLayer :25-44) and transformer/Models.py
(Encoder :14-56, Decoder :59-111, Transformer :114-153) with identical sub-module / parameter names, so a
reference checkpoint of the current schema loads.  `Decoder.forward` follows the evident intent of the
reference (lengths for the masks, positional encoding ADDED, cross-attention residual on
the query) because the file as written raises on every call.

Masks are built on the device from the length vectors (no numpy loop, no per-step H2D copy, no `.item()`
sync — cf. Utils.py:41-70, Embedding.py:22); the key-padding mask stays a stride-0 broadcast view.
The input front-end (Linear -> ReLU -> Dropout -> LayerNorm + positional encoding), the target embedding and the
vocabulary projection run on libst_b200.so as well (st_frontend_*, st_embed_*, st_linear_*), so a
training step launches no PyTorch compute kernel between the batch and the loss; the sub-modules stay ordinary
nn.Linear / nn.LayerNorm / nn.Embedding parameter holders with the reference's names.
"""
import math

import torch
import torch.nn as nn

from . import functional as F
from .transformer.Attention import MultiHeadAttention
from .transformer.SubLayers import PositionwiseFeedForward

PAD = 0  # transformer/Constants.py:1

COMPUTE_DTYPES = {"tf32": torch.float32, "fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16,
                  torch.float32: torch.float32, torch.float16: torch.float16, torch.bfloat16: torch.bfloat16}


# ------------------------------------------------------------------------------------------------ masks (device side)
def key_padding_mask(k_lengths: torch.Tensor, len_q: int, len_k: int) -> torch.Tensor:
    """Same values as Utils.padding_info_mask (Utils.py:41-57): (B, len_q, len_k), True where key >= length;
    an expanded view with stride 0 over the query dimension."""
    ar = torch.arange(len_k, device=k_lengths.device)
    return (ar.unsqueeze(0) >= k_lengths.unsqueeze(1)).unsqueeze(1).expand(-1, len_q, -1)


def subsequent_mask(batch: int, length: int, device) -> torch.Tensor:
    """Same values as Utils.feature_info_mask (Utils.py:60-70): strictly upper triangular."""
    return torch.ones(length, length, dtype=torch.bool, device=device).triu(1).unsqueeze(0).expand(batch, -1, -1)


# ------------------------------------------------------------------------------------------------ Layers.py
class EncoderLayer(nn.Module):
    def __init__(self, d_model, d_inner_hid, n_head, d_k, d_v, dropout=0.1):
        super(EncoderLayer, self).__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner_hid, dropout=dropout)

    def forward(self, inputs, slf_attn_mask=None):
        attn_output, slf_attn_weight = self.slf_attn(inputs, inputs, inputs, mask=slf_attn_mask)
        return self.pos_ffn(attn_output), slf_attn_weight


class DecoderLayer(nn.Module):
    def __init__(self, d_model, d_inner_hid, n_head, d_k, d_v, dropout=0.1):
        super(DecoderLayer, self).__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout)
        self.enc_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout, residual="q")
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner_hid, dropout=dropout)

    def forward(self, inputs, enc_output, slf_attn_mask=None, dec_enc_attn_mask=None):
        slf_attn_output, slf_attn_weight = self.slf_attn(inputs, inputs, inputs, mask=slf_attn_mask)
        sre_attn_output, sre_attn_weight = self.enc_attn(slf_attn_output, enc_output, enc_output, mask=dec_enc_attn_mask)
        return self.pos_ffn(sre_attn_output), (slf_attn_weight, sre_attn_weight)


# ------------------------------------------------------------------------------------------------ Embedding.py
class PositionalEncoding(nn.Module):
    """Sinusoid table (Embedding.py:7-29); forward(n) returns pe[:, :n] (broadcast over the batch)."""

    def __init__(self, dropout, dim, max_len=600):
        super(PositionalEncoding, self).__init__()
        pe = torch.zeros(max_len, dim)
        position = torch.arange(0, max_len).unsqueeze(1)
        div_term = torch.exp((torch.arange(0, dim, 2, dtype=torch.float) * -(math.log(10000.0) / dim)))
        pe[:, 0::2] = torch.sin(position.float() * div_term)
        pe[:, 1::2] = torch.cos(position.float() * div_term)
        self.register_buffer('pe', pe.unsqueeze(0))
        self.dropout = nn.Dropout(p=dropout)   # never applied by the reference either (Embedding.py:21-29)
        self.dim = dim

    def forward(self, time_steps: int):
        return self.pe[:, :time_steps]


# ------------------------------------------------------------------------------------------------ Models.py
class Encoder(nn.Module):
    def __init__(self, input_size, n_max_seq, n_layers=6, n_head=8, d_k=64, d_v=64, d_model=512, d_inner_hid=1024,
                 dropout=0.1, emb_scale=1, compute_dtype=torch.float32, length_masks=True):
        super(Encoder, self).__init__()
        self.compute_dtype, self.length_masks = compute_dtype, length_masks
        self.n_max_seq = n_max_seq
        self.d_model = d_model
        self.emb_scale = emb_scale
        self.position_enc = PositionalEncoding(dropout, d_model, self.n_max_seq)
        self.input_proj = nn.Sequential(nn.Linear(input_size, d_model, bias=True), nn.ReLU(), nn.Dropout(),
                                        nn.LayerNorm(d_model, eps=1e-6))            # Models.py:28-33
        self.layer_stack = nn.ModuleList([EncoderLayer(d_model, d_inner_hid, n_head, d_k, d_v, dropout=dropout)
                                          for _ in range(n_layers)])

    def forward(self, inputs, inputs_length, return_attns=False):
        T = inputs.size(1)
        lin, drop, ln = self.input_proj[0], self.input_proj[2], self.input_proj[3]
        p = drop.p if self.training else 0.0
        enc_output, hidden = F.frontend(inputs, lin.weight, lin.bias, ln.weight, ln.bias, self.position_enc(T)[0],
                                        eps=ln.eps, dropout_p=p, seed=F.next_seed() if p > 0 else 0,
                                        return_hidden=True, out_dtype=self.compute_dtype)   # Models.py:28-33,42-44
        self.last_hidden = hidden if getattr(self, "keep_hidden", False) else None    # test hook (ReLU gate pattern)
        # Models.py:46 — as lengths (the kernels derive the predicate; no mask tensor) or as the reference's mask tensor
        mask = F.LengthMask(inputs_length, T, T) if self.length_masks else key_padding_mask(inputs_length, T, T)
        attns = []
        for layer in self.layer_stack:
            layer.slf_attn.return_attention = bool(return_attns)
            enc_output, a = layer(enc_output, slf_attn_mask=mask)
            if return_attns:
                attns += [a]
        return enc_output, attns


class Decoder(nn.Module):
    def __init__(self, vocab_size, n_max_seq, n_layers=6, n_head=8, d_k=64, d_v=64, d_model=512, d_inner_hid=1024,
                 dropout=0.1, emb_scale=1, compute_dtype=torch.float32, length_masks=True):
        super(Decoder, self).__init__()
        self.compute_dtype, self.length_masks = compute_dtype, length_masks
        self.n_max_seq = n_max_seq
        self.output_dim = vocab_size
        self.d_model = d_model
        self.emb_scale = emb_scale
        self.position_enc = PositionalEncoding(dropout, d_model, self.n_max_seq)
        self.tgt_word_emb = nn.Embedding(vocab_size, d_model, PAD)
        self.layer_stack = nn.ModuleList([DecoderLayer(d_model, d_inner_hid, n_head, d_k, d_v, dropout=dropout)
                                          for _ in range(n_layers)])

    def forward(self, outputs_data, outputs_pos, input_pos, enc_output, return_attns=False):
        B, L = outputs_data.shape
        T = enc_output.size(1)
        dec_output = F.embedding(outputs_data, self.tgt_word_emb.weight, self.position_enc(L)[0], padding_idx=PAD,
                                 out_dtype=self.compute_dtype)                        # Models.py:84-87 (as intended)
        if self.length_masks:   # Models.py:89-97 from the length vectors alone
            slf_mask = F.LengthMask(outputs_pos, L, L, causal=True)
            enc_mask = F.LengthMask(input_pos, L, T)
        else:
            slf_mask = key_padding_mask(outputs_pos, L, L) | subsequent_mask(B, L, outputs_data.device)   # :89-94
            enc_mask = key_padding_mask(input_pos, L, T)                                 # :96-97
        slf_attns, enc_attns = [], []
        for layer in self.layer_stack:
            layer.slf_attn.return_attention = layer.enc_attn.return_attention = bool(return_attns)
            dec_output, (a, c) = layer(dec_output, enc_output, slf_attn_mask=slf_mask, dec_enc_attn_mask=enc_mask)
            if return_attns:
                slf_attns += [a]
                enc_attns += [c]
        return dec_output, slf_attns, enc_attns


class Transformer(nn.Module):
    """Models.py:114-153.  `config` needs the attributes the reference reads (Models.py:120-143)."""

    def __init__(self, config):
        super(Transformer, self).__init__()
        self.return_attns = bool(getattr(config, "return_attns", False))
        # compute_dtype: "tf32" (fp32 storage, TF32 tensor-core operands — BASELINE.json configs[1]), "fp16" or "bf16"
        # (16-bit activations and tensor-core operands, fp32 parameters / statistics / accumulation — configs[2])
        self.compute_dtype = COMPUTE_DTYPES[getattr(config, "compute_dtype", None) or "tf32"]
        common = dict(n_head=config.n_heads, d_k=config.d_k, d_v=config.d_v, d_model=config.d_model,
                      d_inner_hid=config.d_inner_hid, dropout=config.dropout, emb_scale=getattr(config, "emb_scale", 1),
                      compute_dtype=self.compute_dtype,
                      length_masks=bool(getattr(config, "length_masks", True)))
        self.encoder = Encoder(input_size=config.feature_dim, n_max_seq=config.max_inputs_length,
                               n_layers=config.num_enc_layer, **common)
        self.decoder = Decoder(vocab_size=config.vocab_size, n_max_seq=config.max_target_length,
                               n_layers=config.num_dec_layer, **common)
        self.tgt_word_proj = nn.Linear(config.d_model, config.vocab_size, bias=False)

    def forward(self, inputs, inputs_pos, targets=None, targets_pos=None):
        enc_output, enc_slf_attn = self.encoder(inputs, inputs_pos, self.return_attns)
        dec_output, dec_slf_attn, dec_enc_attn = self.decoder(targets, targets_pos, inputs_pos, enc_output,
                                                              self.return_attns)
        logits = F.linear(dec_output, self.tgt_word_proj.weight, self.tgt_word_proj.bias)     # Models.py:151
        return logits, (enc_slf_attn, dec_slf_attn, dec_enc_attn)


class ModelConfig(dict):
    """Attribute-style config (like Utils.AttrDict, Utils.py:9-22, but a missing key raises)."""

    def __getattr__(self, item):
        try:
            return self[item]
        except KeyError as e:
            raise AttributeError(item) from e


def headline_config(**over) -> ModelConfig:
    """BASELINE.json configs[1]: 6+6 layers, d_model 512, 8 heads, d_ff 2048, 80-dim fbank, V=4337."""
    cfg = ModelConfig(feature_dim=80, vocab_size=4337, max_inputs_length=2048, max_target_length=64, d_model=512,
                      n_heads=8, d_k=64, d_v=64, d_inner_hid=2048, num_enc_layer=6, num_dec_layer=6, dropout=0.1,
                      emb_scale=1, return_attns=False)
    cfg.update(over)
    return cfg


def init_parameters(model: nn.Module) -> None:
    """Utils.init_parameters (Utils.py:101-104): xavier_normal_ on every parameter with dim >= 2."""
    for _, p in model.named_parameters():
        if p.dim() >= 2:
            nn.init.xavier_normal_(p)
