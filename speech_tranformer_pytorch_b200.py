"""Import shim: the package directory is named `speech-tranformer-pytorch_b200` (with hyphens, as the
project layout requires), which Python cannot import by name.  `import speech_tranformer_pytorch_b200`
loads that directory as a regular package under this module's name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "speech-tranformer-pytorch_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
