"""Make the reference's own composition code (transformer/Layers.py, Models.py, train.py) pick up the
B200 modules: `from transformer.Attention import MultiHeadAttention` (Layers.py:3-4) resolves through
sys.modules, so registering our modules under those names before `transformer.Layers` is imported is
all the integration the reference needs."""
import sys

_NAMES = ("Attention", "SubLayers", "Loss")
_saved = {}


def install(package: str = "transformer") -> None:
    from . import transformer as ours
    for n in _NAMES:
        key = f"{package}.{n}"
        _saved.setdefault(key, sys.modules.get(key))
        sys.modules[key] = getattr(ours, n)
        parent = sys.modules.get(package)
        if parent is not None:
            setattr(parent, n, getattr(ours, n))


def uninstall(package: str = "transformer") -> None:
    for n in _NAMES:
        key = f"{package}.{n}"
        old = _saved.pop(key, None)
        if old is None:
            sys.modules.pop(key, None)
        else:
            sys.modules[key] = old
