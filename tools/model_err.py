"""Whole-model logits error against the reference fixture, with TF32-rounded (default) or exact-fp32 (ST_ROUND_OUT=0)
residual streams — the measurement behind TOL_MODEL in tests/test_gpu_model.py.  python tools/model_err.py"""
import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import speech_tranformer_pytorch_b200 as stb
from helpers import golden, relerr, t
import test_gpu_model as tm
g = golden("transformer_small"); V = 31
net = tm._small_model(stb, g)
batch = [t(g[k], "cuda:0") for k in ("inputs", "in_len", "targets", "tgt_len")]
logits, _ = net(*batch)
print("ROUND_OUT", stb.functional.ROUND_OUT, "logits relerr", relerr(logits, g["logits"]))
