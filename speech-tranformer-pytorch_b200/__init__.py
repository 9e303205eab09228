"""speech-tranformer-pytorch_b200 — the Speech-Transformer hot path (attention, position-wise FFN,
residual+LayerNorm, label-smoothed CE) as hand-written sm_100a CUDA behind a C ABI, wrapped as
drop-in nn.Modules with the reference's interface.

    import speech_tranformer_pytorch_b200 as stb          # via the repo-root shim
    stb.install()                                         # `transformer.Attention` etc. now resolve here
    from transformer.Layers import EncoderLayer           # the reference's own composition code
"""
from . import _lib, functional, transformer  # noqa: F401
from .build import build  # noqa: F401
from .install import install, uninstall  # noqa: F401
from .transformer import (CrossEntropyLoss, CTCLoss, JointCTCAttentionLoss, LabelSmoothingLoss, MultiHeadAttention,  # noqa: F401
                          PositionwiseFeedForward, ScaledDotProductAttention)

__version__ = "0.1.0"
