"""CPU ORACLE (test / baseline infrastructure only) — whole-model port used when oracle/_ref (the scratch
copy of the reference package) is not available.

Restates transformer/Models.py (Encoder :14-56, Decoder :59-111 as intended, see SURVEY.md §3.2,
Transformer :114-153) functionally on top of oracle/st_oracle.py, and train.py:37-46 (zero_grad, forward,
loss, backward, clip_grad_norm_, Noam-Adam).  Dropout is not modelled (p = 0), which the caller reports.
"""
import math

import torch

from . import st_oracle as O


def sinusoid(max_len: int, dim: int) -> torch.Tensor:
    """Embedding.py:8-17."""
    pe = torch.zeros(max_len, dim)
    position = torch.arange(0, max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=torch.float) * -(math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position.float() * div_term)
    pe[:, 1::2] = torch.cos(position.float() * div_term)
    return pe


def init_params(cfg: dict, seed: int = 2018) -> dict:
    g = torch.Generator().manual_seed(seed)
    d, f, F, V = cfg["d_model"], cfg["d_inner_hid"], cfg["feature_dim"], cfg["vocab_size"]
    P = {}

    def lin(name, out, inp, bias=True):
        P[name + ".weight"] = torch.nn.init.xavier_normal_(torch.empty(out, inp), generator=g)
        if bias:
            P[name + ".bias"] = torch.zeros(out)

    def ln(name):
        P[name + ".weight"], P[name + ".bias"] = torch.ones(d), torch.zeros(d)

    def mha(prefix):
        for n in ("linear_q", "linear_k", "linear_v", "output_linear"):
            lin(prefix + n, d, d)
        ln(prefix + "layernorm")

    def ffn(prefix):
        lin(prefix + "fc1", f, d)
        lin(prefix + "fc2", d, f)
        ln(prefix + "layernorm")

    lin("encoder.input_proj.0", d, F)
    ln("encoder.input_proj.3")
    for i in range(cfg["num_enc_layer"]):
        mha(f"encoder.layer_stack.{i}.slf_attn.")
        ffn(f"encoder.layer_stack.{i}.pos_ffn.")
    P["decoder.tgt_word_emb.weight"] = torch.nn.init.xavier_normal_(torch.empty(V, d), generator=g)
    for i in range(cfg["num_dec_layer"]):
        mha(f"decoder.layer_stack.{i}.slf_attn.")
        mha(f"decoder.layer_stack.{i}.enc_attn.")
        ffn(f"decoder.layer_stack.{i}.pos_ffn.")
    lin("tgt_word_proj", V, d, bias=False)
    return {k: v.requires_grad_() for k, v in P.items()}


def forward(P: dict, cfg: dict, inputs, in_len, targets, tgt_len, gate_fn=None):
    """`gate_fn(name, pre_activation) -> 0/1 gate or None` lets a test pin the active set of every ReLU (the
    front-end's, "frontend", and each layer's FFN, "encoder.layer_stack.i" / "decoder.layer_stack.i") to the one a
    reduced-precision forward used — gradients of a piecewise-linear function are only comparable for a fixed
    active set (tests/helpers.relu_gate_from_cuda)."""
    d, H = cfg["d_model"], cfg["n_heads"]
    T, L = inputs.size(1), targets.size(1)
    dt = P["tgt_word_proj.weight"].dtype
    pe = sinusoid(max(T, L), d).to(dt)
    inputs = inputs.to(dt)

    def relu(name, pre):
        gate = gate_fn(name, pre) if gate_fn is not None else None
        return torch.relu(pre) if gate is None else pre * gate.to(pre.dtype)

    def ffn(name, x, prm):
        gate = gate_fn(name, O.ffn_preactivation(x, prm)) if gate_fn is not None else None
        return O.positionwise_ffn(x, prm, gate=gate)

    x = relu("frontend", inputs @ P["encoder.input_proj.0.weight"].t() + P["encoder.input_proj.0.bias"])  # Models.py:28-33
    x = O.add_layer_norm(x, None, P["encoder.input_proj.3.weight"], P["encoder.input_proj.3.bias"]) + pe[:T]
    enc_mask = O.padding_info_mask(in_len, in_len).bool()                                                # Models.py:46
    sub = lambda pre: {k[len(pre):]: v for k, v in P.items() if k.startswith(pre)}
    for i in range(cfg["num_enc_layer"]):
        name = f"encoder.layer_stack.{i}"
        a, _ = O.multi_head_attention(x, x, x, enc_mask, sub(name + ".slf_attn."), H)                    # Layers.py:18-22
        x = ffn(name, a, sub(name + ".pos_ffn."))
    y = P["decoder.tgt_word_emb.weight"][targets] + pe[:L]                                               # Models.py:84-87
    slf_mask = O.decoder_self_mask(tgt_len)                                                              # Models.py:89-94
    cross_mask = O.padding_info_mask(tgt_len, in_len).bool()                                             # Models.py:96-97
    for i in range(cfg["num_dec_layer"]):
        name = f"decoder.layer_stack.{i}"
        a, _ = O.multi_head_attention(y, y, y, slf_mask, sub(name + ".slf_attn."), H)                    # Layers.py:37-44
        c, _ = O.multi_head_attention(a, x, x, cross_mask, sub(name + ".enc_attn."), H, residual="q")
        y = ffn(name, c, sub(name + ".pos_ffn."))
    return y @ P["tgt_word_proj.weight"].t()                                                             # Models.py:151


def make_train_step(cfg: dict, inputs, targets, in_len, tgt_len, truth, max_grad_norm: float = 5.0, warmup: int = 12000):
    P = init_params(cfg)
    V, d = cfg["vocab_size"], cfg["d_model"]
    one_hot = O.smoothing_one_hot(0.1, V, 0)
    weight = torch.ones(V)
    opt = torch.optim.Adam(list(P.values()), lr=0.0, betas=(0.9, 0.98), eps=1e-9)                        # Optim.py:9-14
    state = {"step": 0}

    def step():
        state["step"] += 1
        s = state["step"]
        opt.zero_grad()
        logits = forward(P, cfg, inputs, in_len, targets, tgt_len)
        loss = O.label_smoothing_loss(logits.reshape(-1, V), truth.reshape(-1), one_hot, weight, 0.1, 0, True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(P.values()), max_grad_norm)
        for gparam in opt.param_groups:
            gparam["lr"] = d ** -0.5 * min(s ** -0.5, s * warmup ** -1.5)                                # Optim.py:36-45
        opt.step()
        return float(loss)

    return step
