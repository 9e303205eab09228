"""CPU-side checks (no GPU): the C-ABI library builds, loads and exports every symbol include/st_b200.h
declares; the ctypes structures match the header; the host-side mirror keeps the reference's interface
(class names, constructor signatures, parameter names, error behaviour); there is no CPU fallback."""
import ctypes
import inspect
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "st_b200.h")


@pytest.fixture(scope="module")
def stb(lib_path):
    import speech_tranformer_pytorch_b200 as m
    return m


def _declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(st_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(stb, lib_path):
    names = _declared_symbols()
    assert len(names) >= 25
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(stb._lib.SIGNATURES) == names, "the ctypes table and the header must list the same entry points"


def test_header_is_plain_c(tmp_path):
    """The boundary must be consumable from C: compile a translation unit that only includes the header."""
    src = tmp_path / "t.c"
    src.write_text('#include "st_b200.h"\nint main(void) { st_gemm_epilogue e; (void)e; return sizeof(st_mha_args) > 0 ? 0 : 1; }\n')
    rc = os.system(f"gcc -std=c99 -Wall -Werror -I {os.path.join(ROOT, 'include')} -c {src} -o {tmp_path / 't.o'}")
    assert rc == 0


def test_ctypes_struct_layout_matches_header(stb, tmp_path):
    """sizeof / offsetof of every argument struct, computed by the C compiler, equals the ctypes mirror."""
    L = stb._lib
    pairs = {"st_gemm_epilogue": L.GemmEpilogue, "st_attn_args": L.AttnArgs, "st_attn_bwd_args": L.AttnBwdArgs,
             "st_mha_args": L.MhaArgs, "st_mha_bwd_args": L.MhaBwdArgs, "st_ffn_args": L.FfnArgs,
             "st_ffn_bwd_args": L.FfnBwdArgs, "st_adam_args": L.AdamArgs, "st_frontend_args": L.FrontendArgs,
             "st_frontend_bwd_args": L.FrontendBwdArgs, "st_linear_args": L.LinearArgs, "st_linear_bwd_args": L.LinearBwdArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "st_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    assert os.system(f"gcc -std=c99 -I {os.path.join(ROOT, 'include')} {src} -o {exe}") == 0
    got = dict(l.split() for l in os.popen(str(exe)).read().splitlines())
    for cname, cls in pairs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_load_and_cpu_only_calls(stb):
    lib = stb._lib.load()
    assert lib.st_version() >= 100
    assert lib.st_selftest_count() >= 15
    assert lib.st_launch_count() == 0                      # nothing has been launched in this process
    assert lib.st_set_option(b"no_such_option", 1) != 0
    assert b"unknown option" in lib.st_last_error()
    assert lib.st_ffn_saved_floats(10, 64, 128, 1) < lib.st_ffn_saved_floats(10, 64, 128, 0)
    assert lib.st_mha_saved_floats(2, 5, 5, 2, 64, 1, 1, 0) > 0


def test_no_cpu_fallback(stb):
    m = stb.MultiHeadAttention(2, 64, 32, 32)
    x = torch.randn(1, 4, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, x, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        stb.PositionwiseFeedForward(64, 128)(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        stb.LabelSmoothingLoss(0.1, 10, weight=torch.ones(10))(torch.randn(3, 10), torch.tensor([1, 2, 3]))


def test_missing_library_fails_loudly(stb, monkeypatch):
    monkeypatch.setattr(stb._lib, "_lib", None)
    monkeypatch.setattr(stb._lib, "LIB_PATH", "/nonexistent/libst_b200.so")
    with pytest.raises(stb._lib.StError, match="no CPU or PyTorch fallback"):
        stb._lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "speech-tranformer-pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


# ------------------------------------------------------------------------------------------------ interface mirror
def test_module_interfaces_match_reference(stb):
    """Constructor signatures and parameter names of transformer/{Attention,SubLayers,Loss}.py (reference lines cited
    in the modules' docstrings); the extra keyword arguments are appended with defaults."""
    def params(fn):
        return list(inspect.signature(fn).parameters)
    assert params(stb.MultiHeadAttention.__init__)[:6] == ["self", "n_head", "d_model", "d_k", "d_v", "dropout"]
    assert params(stb.MultiHeadAttention.forward) == ["self", "q", "k", "v", "mask"]
    assert params(stb.ScaledDotProductAttention.__init__) == ["self", "d_k", "dropout"]
    assert params(stb.PositionwiseFeedForward.__init__) == ["self", "d_model", "d_ff", "dropout"]
    assert params(stb.LabelSmoothingLoss.__init__) == ["self", "label_smoothing", "vocab_size", "weight", "size_average", "ignore_index"]
    assert params(stb.CrossEntropyLoss.__init__) == ["self", "weight", "size_average"]
    sd = stb.MultiHeadAttention(8, 512, 64, 64).state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "linear_q.weight": (512, 512), "linear_q.bias": (512,), "linear_k.weight": (512, 512), "linear_k.bias": (512,),
        "linear_v.weight": (512, 512), "linear_v.bias": (512,), "output_linear.weight": (512, 512),
        "output_linear.bias": (512,), "layernorm.weight": (512,), "layernorm.bias": (512,)}
    sd = stb.PositionwiseFeedForward(512, 2048).state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "fc1.weight": (2048, 512), "fc1.bias": (2048,), "fc2.weight": (512, 2048), "fc2.bias": (512,),
        "layernorm.weight": (512,), "layernorm.bias": (512,)}
    crit = stb.LabelSmoothingLoss(0.1, 30, weight=torch.ones(30), ignore_index=0)
    assert tuple(crit.one_hot.shape) == (1, 30) and crit.one_hot[0, 0] == 0 and "one_hot" in crit.state_dict()
    assert stb.LabelSmoothingLoss(0.1, 30, ignore_index=3).one_hot[0, 3] != 0      # `if not ignore_index` quirk (Loss.py:20)
    with pytest.raises(AssertionError):
        stb.MultiHeadAttention(3, 64, 32, 32)                                      # Attention.py:45
    with pytest.raises(AssertionError):
        stb.LabelSmoothingLoss(1.5, 30)                                            # Loss.py:14


def test_install_redirects_reference_imports(stb):
    """`from transformer.Attention import MultiHeadAttention` (Layers.py:3-4) resolves to the B200 modules."""
    saved = {k: v for k, v in sys.modules.items() if k == "transformer" or k.startswith("transformer.")}
    for k in saved:
        del sys.modules[k]
    try:
        stb.install()
        from transformer.Attention import MultiHeadAttention
        from transformer.SubLayers import PositionwiseFeedForward
        from transformer.Loss import LabelSmoothingLoss
        assert MultiHeadAttention is stb.MultiHeadAttention
        assert PositionwiseFeedForward is stb.PositionwiseFeedForward and LabelSmoothingLoss is stb.LabelSmoothingLoss
    finally:
        stb.uninstall()
        for k in [k for k in sys.modules if k == "transformer" or k.startswith("transformer.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_model_assembly_and_device_masks(stb):
    """Model assembly mirrors Models.py names; device-side masks equal the oracle's (= the reference's) byte for byte."""
    from oracle import st_oracle as O
    from speech_tranformer_pytorch_b200 import model as M
    lens_q, lens_k = torch.tensor([4, 2, 3]), torch.tensor([7, 5, 2])
    ref = O.padding_info_mask(lens_q, lens_k)
    got = M.key_padding_mask(lens_k, 4, 7)
    assert got.stride(1) == 0 and torch.equal(got, ref.bool())
    lens = torch.tensor([5, 3])
    assert torch.equal(M.subsequent_mask(2, 5, "cpu"), O.feature_info_mask(lens).bool())
    assert torch.equal(M.key_padding_mask(lens, 5, 5) | M.subsequent_mask(2, 5, "cpu"), O.decoder_self_mask(lens))
    net = M.Transformer(M.headline_config(num_enc_layer=1, num_dec_layer=1))
    names = set(net.state_dict())
    for n in ("encoder.input_proj.0.weight", "encoder.input_proj.3.bias", "encoder.position_enc.pe",
              "encoder.layer_stack.0.slf_attn.linear_q.weight", "encoder.layer_stack.0.pos_ffn.fc1.weight",
              "decoder.tgt_word_emb.weight", "decoder.layer_stack.0.enc_attn.output_linear.bias", "tgt_word_proj.weight"):
        assert n in names, n
    assert sum(p.numel() for p in M.Transformer(M.headline_config()).parameters()) == 48_622_080   # 48.6 M


def test_synthetic_batch_matches_oracle_generator(stb):
    from oracle import st_oracle as O
    from speech_tranformer_pytorch_b200 import data as D
    a = D.synthetic_batch(3, 40, 12, 80, 100, seed=5, fixed_len=False, t_min=10, l_min=3)
    b = O.synthetic_batch(3, 40, 12, 80, 100, seed=5, fixed_len=False, t_min=10, l_min=3)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_python_sources_reference_only_defined_names():
    """The GPU-only code paths (autograd backward functions, trainers, bench) cannot run in the CPU suite; at least make
    sure every name they load is defined somewhere in their module (a deleted helper shows up here, not on the GPU box)."""
    import ast
    import builtins
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "speech-tranformer-pytorch_b200", "**", "*.py"), recursive=True)
    files += [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")] + glob.glob(os.path.join(root, "tools", "*.py"))
    bad = []
    for path in files:
        tree = ast.parse(open(path).read())
        known = set(dir(builtins)) | {"__file__", "__name__"}
        for n in ast.walk(tree):
            if isinstance(n, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
                known.add(n.name)
            elif isinstance(n, (ast.Import, ast.ImportFrom)):
                known.update((a.asname or a.name).split(".")[0] for a in n.names)
            elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
                known.add(n.id)
            elif isinstance(n, ast.arg):
                known.add(n.arg)
            elif isinstance(n, ast.ExceptHandler) and n.name:
                known.add(n.name)
        bad += [f"{os.path.relpath(path, root)}:{n.lineno} {n.id}" for n in ast.walk(tree)
                if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in known]
    assert not bad, bad


def test_token_constants_match_reference():
    """decode.BOS/EOS, data.*, the oracle's decode port and the reference's Constants.py (PAD 0, UNK 1, BOS 2, EOS 3) agree:
    a checkpoint trained with BOS = 2 as the start symbol must be decoded from the same symbol."""
    import importlib
    import speech_tranformer_pytorch_b200 as stb
    data = importlib.import_module(stb.__name__ + ".data")
    decode = importlib.import_module(stb.__name__ + ".decode")
    from oracle import decode_port
    want = dict(PAD=0, UNK=1, BOS=2, EOS=3)
    assert (data.PAD, data.UNK, data.BOS, data.EOS) == (0, 1, 2, 3)
    assert (decode.PAD, decode.BOS, decode.EOS) == (want["PAD"], want["BOS"], want["EOS"])
    assert (decode_port.PAD, decode_port.UNK, decode_port.BOS, decode_port.EOS) == (0, 1, 2, 3)
    ref = os.path.join(ROOT, "oracle", "_ref", "transformer", "Constants.py")
    if os.path.exists(ref):
        ns = {}
        exec(open(ref).read(), ns)
        assert {k: ns[k] for k in want} == want


def test_st_options_environment_is_applied_at_load(lib_path):
    """ST_OPTIONS="name=value,..." sets library options when the library is loaded; unknown names fail loudly."""
    import subprocess
    code = ("import speech_tranformer_pytorch_b200 as m; lib = m._lib.load(); "
            "import ctypes; print('loaded')")
    env = dict(os.environ, ST_OPTIONS="gemm_rows16=0,side_streams=0")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "loaded" in r.stdout, r.stderr[-500:]
    env = dict(os.environ, ST_OPTIONS="no_such_option=1")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "unknown option" in r.stderr
