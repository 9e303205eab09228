"""Generate tests/golden/transformer_small.npz: the reference's WHOLE model (oracle/_ref — the scratch copy of
/root/reference/transformer with the compatibility patch of oracle/make_ref.py) run in eval mode on a seeded
ragged batch, with the label-smoothed loss of train.py:120 and every parameter gradient.  Build container only.

    python oracle/make_ref.py && python oracle/make_golden_model.py

Config = BASELINE.json configs[0] shape (2+2 layers, d_model 64, 2 heads, d_ff 128) on 80-dim features.
Pins oracle/model_port.py (tests/test_oracle_golden.py) and the CUDA model (tests/test_gpu_model.py).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
for name in ("editdistance", "matplotlib", "matplotlib.pyplot", "tensorboardX"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, os.path.join(HERE, "_ref"))

from transformer.Models import Transformer  # noqa: E402
from transformer.Loss import LabelSmoothingLoss  # noqa: E402
from transformer.Utils import AttrDict  # noqa: E402
from oracle import st_oracle as O  # noqa: E402

CFG = dict(feature_dim=80, vocab_size=31, max_inputs_length=64, max_target_length=16, d_model=64, n_heads=2, d_k=32,
           d_v=32, d_inner_hid=128, num_enc_layer=2, num_dec_layer=2, dropout=0.1, emb_scale=1, return_attns=False)


def main():
    torch.manual_seed(2018)
    V = CFG["vocab_size"]
    model = Transformer(AttrDict(CFG))
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2:
                torch.nn.init.xavier_normal_(p, generator=gen)
            elif name.endswith("layernorm.weight") or name.endswith("input_proj.3.weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=gen))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=gen))
    model.eval()
    inputs, targets, in_len, tgt_len, truth = O.synthetic_batch(3, 37, 9, 80, V, seed=11, fixed_len=False, t_min=20, l_min=4)
    crit = LabelSmoothingLoss(0.1, V, weight=torch.ones(V), size_average=True, ignore_index=0)
    logits, _ = model(inputs, in_len, targets, tgt_len)
    loss = crit(logits.contiguous().view(-1, V), truth.contiguous().view(-1))
    loss.backward()
    out = {"inputs": inputs.numpy(), "targets": targets.numpy(), "in_len": in_len.numpy(), "tgt_len": tgt_len.numpy(),
           "truth": truth.numpy(), "logits": logits.detach().numpy(), "loss": loss.detach().numpy()}
    for k, v in model.state_dict().items():
        out["p." + k] = v.detach().numpy()
    for k, p in model.named_parameters():
        out["g." + k] = p.grad.detach().numpy()
    path = os.path.join(ROOT, "tests", "golden", "transformer_small.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, loss {float(loss):.6f}")


if __name__ == "__main__":
    main()
