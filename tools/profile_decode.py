"""Per-kernel-class CUDA-event times of one beam-decode run (eager launches) of the 6+6 x 512 x 8 model.
python tools/profile_decode.py [--batch 32] [--frames 1000] [--beam 10] [--steps 50]"""
import argparse, collections, ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
from speech_tranformer_pytorch_b200 import data as sdata, decode, model as smodel
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32); ap.add_argument("--frames", type=int, default=1000)
ap.add_argument("--beam", type=int, default=10); ap.add_argument("--steps", type=int, default=50)
a = ap.parse_args()
dev = torch.device("cuda", 0); lib = stb._lib.load()
torch.manual_seed(2018)
net = smodel.Transformer(smodel.headline_config()); smodel.init_parameters(net); net = net.to(dev).eval()
inputs, _, in_len, _, _ = [t.to(dev) for t in sdata.synthetic_batch(a.batch, a.frames, 50, 80, 4337)]
decode.beam_search(net, inputs, in_len, beam=a.beam, max_len=a.steps); torch.cuda.synchronize()
dec = decode.IncrementalDecoder(net, max_len=a.steps)
dec.start(inputs, in_len, a.beam); torch.cuda.synchronize()
tok = torch.full((a.batch * a.beam,), 1, dtype=torch.int64, device=dev)
lib.st_profile_reset(); lib.st_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps): dec.step(tok)
e1.record(); torch.cuda.synchronize(); lib.st_profile_enable(0)
print(f"{a.steps} positions: {e0.elapsed_time(e1):.1f} ms wall (with per-launch events)")
tot = 0
for c in range(lib.st_profile_classes()):
    t, w, n = C.c_double(), C.c_double(), C.c_int64()
    lib.st_profile_read(c, C.byref(t), C.byref(w), C.byref(n))
    if n.value:
        print(f"  {lib.st_profile_class_name(c).decode():14s} {n.value:5d} launches  {t.value:8.2f} ms  avg {t.value / n.value * 1e3:7.1f} us")
        tot += t.value
print(f"  library kernels total {tot:.1f} ms")
