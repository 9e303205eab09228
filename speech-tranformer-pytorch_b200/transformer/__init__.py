"""Drop-in mirror of the reference's `transformer` package for the hot path only:
`transformer.Attention`, `transformer.SubLayers`, `transformer.Loss` (same class names, constructor
signatures, parameter names and forward contracts), backed by libst_b200.so."""
from . import Attention, Loss, SubLayers  # noqa: F401
from .Attention import MultiHeadAttention, ScaledDotProductAttention  # noqa: F401
from .Loss import CrossEntropyLoss, CTCLoss, JointCTCAttentionLoss, LabelSmoothingLoss  # noqa: F401
from .SubLayers import PositionwiseFeedForward  # noqa: F401
