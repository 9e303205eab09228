"""B200 drop-in for the reference's transformer/SubLayers.py (PositionwiseFeedForward, :9-28)."""
import torch.nn as nn
import torch.nn.init as init

from .. import functional as F

__all__ = ["PositionwiseFeedForward"]


class PositionwiseFeedForward(nn.Module):
    """dropout2(LayerNorm(x + fc2(dropout1(relu(fc1(x)))))) — SubLayers.py:24-28, one fused operator."""

    def __init__(self, d_model, d_ff, dropout=0.1):
        super().__init__()
        # member names, registration order, shapes and the order in which the initialisers draw random numbers are the
        # reference's (SubLayers.py:13-22): state_dict / init_parameters / same-seed construction stay interchangeable
        projections = {"fc1": (d_model, d_ff), "fc2": (d_ff, d_model)}
        for name, (fan_in, fan_out) in projections.items():
            setattr(self, name, nn.Linear(fan_in, fan_out))
        self.relu = nn.ReLU()                                   # parameter-free members kept for attribute compatibility
        self.dropout1, self.dropout2 = nn.Dropout(dropout), nn.Dropout(dropout)
        self.layernorm = nn.LayerNorm(d_model, eps=1e-6)
        for name in projections:
            init.xavier_normal_(getattr(self, name).weight)
        self.keep_hidden = False     # test hook: keep the hidden activation of the last forward in .last_hidden
        self.last_hidden = None

    def forward(self, inputs):
        p = self.dropout1.p if self.training else 0.0
        out, hidden = F.positionwise_ffn(inputs, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias,
                                         self.layernorm.weight, self.layernorm.bias, eps=self.layernorm.eps,
                                         dropout_p=p, seed=F.next_seed() if p > 0 else 0, round_out=F.ROUND_OUT,
                                         return_hidden=True)
        self.last_hidden = hidden if self.keep_hidden else None
        return out
