"""BASELINE.json configs[4]: beam decode (width 10) of the 6+6 x 512 x 8 model with K/V reuse on 1 x B200.
python tools/bench_decode.py [--batch 32] [--frames 1000] [--beam 10] [--steps 50]
Reports the encoder + cross-K/V setup time, the per-position decode step time and utterances / tokens per second.
Random-init weights never emit EOS, so exactly `steps` positions are decoded (worst case for a 50-symbol target)."""
import argparse, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
from speech_tranformer_pytorch_b200 import data as sdata, decode, model as smodel
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32); ap.add_argument("--frames", type=int, default=1000)
ap.add_argument("--graphs", action="store_true"); ap.add_argument("--beam", type=int, default=10); ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--opt", action="append", default=[], help="library option name=value (st_set_option), repeatable")
a = ap.parse_args()
for kv in a.opt:
    k, v = kv.split("=")
    stb._lib.check(stb._lib.load().st_set_option(k.encode(), int(v)))
dev = torch.device("cuda", 0)
torch.manual_seed(2018)
net = smodel.Transformer(smodel.headline_config()); smodel.init_parameters(net); net = net.to(dev).eval()
inputs, _, in_len, _, _ = [t.to(dev) for t in sdata.synthetic_batch(a.batch, a.frames, 50, 80, 4337)]
lib = stb._lib.load()
ap_graph = "--graphs" in sys.argv
dec_obj = decode.IncrementalDecoder(net, max_len=a.steps, use_graphs=True)
def run():
    return decode.beam_search(net, inputs, in_len, beam=a.beam, max_len=a.steps, decoder=dec_obj)
run(); run(); run(); torch.cuda.synchronize()      # eager pass, capturing pass, first replay
l0 = lib.st_launch_count()
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record(); hyps, scores = run(); e[1].record(); torch.cuda.synchronize()
ms = e[0].elapsed_time(e[1])
launches = lib.st_launch_count() - l0
# split: setup (encoder + cross K/V) vs steps
dec = decode.IncrementalDecoder(net, max_len=a.steps)
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); dec.start(inputs, in_len, a.beam); e[1].record()
tok = torch.full((a.batch * a.beam,), 1, dtype=torch.int64, device=dev)
for _ in range(a.steps): dec.step(tok)
e[2].record(); torch.cuda.synchronize()
print(json.dumps({"workload": f"beam decode width {a.beam}, 6+6 x 512 x 8, B={a.batch}, T={a.frames}, {a.steps} positions",
                  "total_ms": ms, "utterances_per_s": a.batch / (ms * 1e-3), "tokens_per_s": a.batch * a.steps / (ms * 1e-3),
                  "setup_ms": e[0].elapsed_time(e[1]), "ms_per_position": e[1].elapsed_time(e[2]) / a.steps,
                  "library_launches": int(launches), "cuda_graphs": True,
                  "note": "total_ms: persistent decoder replaying per-position CUDA graphs; setup_ms / ms_per_position: eager launches"}))
