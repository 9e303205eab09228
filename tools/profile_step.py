"""One profiled training step of the headline model: per-(kernel class, work size) launch statistics.
python tools/profile_step.py [--layers 6] [--batch 32] [--frames 1000] [--dropout 0.1] [--csv out.csv]"""
import argparse, collections, ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
from speech_tranformer_pytorch_b200 import data as sdata, model as smodel, parallel as spar
ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=6); ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--frames", type=int, default=1000); ap.add_argument("--dropout", type=float, default=0.1)
ap.add_argument("--csv", default="gpurun_out/profile_step.csv"); ap.add_argument("--dtype", default="fp32", choices=["fp32", "tf32", "fp16", "bf16"])  # as bench.py --dtype
a = ap.parse_args()
lib = stb._lib.load(); dev = torch.device("cuda", 0); V = 4337
torch.manual_seed(2018)
stb.functional.set_fp32_engine("tf32" if a.dtype == "tf32" else "fp16")
net = smodel.Transformer(smodel.headline_config(num_enc_layer=a.layers, num_dec_layer=a.layers, dropout=a.dropout, compute_dtype=a.dtype))
smodel.init_parameters(net); net = net.to(dev).train()
crit = stb.LabelSmoothingLoss(0.1, V, weight=torch.ones(V, device=dev), ignore_index=0).to(dev)
tr = spar.DataParallelTrainer(net, d_model=512, compute_dtype=smodel.COMPUTE_DTYPES[a.dtype])
batch = [t.to(dev) for t in sdata.synthetic_batch(a.batch, a.frames, 50, 80, V)]
def step():
    inputs, targets, il, tl, truth = batch
    return tr.train_step(lambda: crit(net(inputs, il, targets, tl)[0].view(-1, V), truth.view(-1)))
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print(f"unprofiled step: {e0.elapsed_time(e1):.2f} ms")
lib.st_set_option(b"side_streams", 0)     # per-launch event intervals must not contain another stream's kernels
lib.st_profile_reset(); lib.st_profile_enable(1); step(); torch.cuda.synchronize(); lib.st_profile_enable(0)
os.makedirs(os.path.dirname(a.csv) or ".", exist_ok=True)
lib.st_profile_dump(a.csv.encode())
names = [lib.st_profile_class_name(i).decode() for i in range(lib.st_profile_classes())]
groups = collections.OrderedDict()
for line in open(a.csv).read().splitlines()[1:]:
    _, cls, work, ms, tag = line.split(","); k = (int(cls), float(work), int(tag)); g = groups.setdefault(k, [0, 0.0]); g[0] += 1; g[1] += float(ms)
tot = sum(g[1] for g in groups.values())
print(f"profiled kernels total {tot:.2f} ms")
def tagstr(t):
    if not t: return ""
    bn, var, fl, mode = t & 15, (t >> 4) & 15, (t >> 8) & 15, (t >> 12) & 3
    K, N, M = ((t >> 16) & 0x3FFFF) * 8, (t >> 34) & 0x3FFF, (t >> 48) * 8
    return f" M{M} N{N} K{K} {'NT NN TN'.split()[mode]} fl{fl} var{var} bn{bn * 64}"
for (cls, work, tag), (n, ms) in sorted(groups.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{names[cls]:15s}{tagstr(tag):44s} work {work:14.4g} x{n:3d}  total {ms:7.3f} ms  avg {ms / n * 1e3:8.1f} us  rate {work * n / ms / 1e9:8.1f} G/s*1e3")

# ---- algorithmic HBM bytes of the GEMM class (operands read once, output written once, aux operand read once)
gb = 0.0
for (cls, work, tag), (n, ms) in groups.items():
    if names[cls] != "gemm_tf32" or not tag: continue
    fl, mode = (tag >> 8) & 15, (tag >> 12) & 3
    K, N, M = ((tag >> 16) & 0x3FFFF) * 8, (tag >> 34) & 0x3FFF, (tag >> 48) * 8
    s_in = 4 if a.dtype == "tf32" else 2
    if a.dtype == "tf32" or fl == 5 or (fl == 2 and mode == 0) or N >= 4336: s_out = 4
    else: s_out = 2
    if K <= 80 or (fl == 5 and N == 80): s_in, s_out = 4, 4          # the 80-wide front-end GEMMs stay TF32
    if a.dtype == "fp32":                                            # fp32 model, fp16 operands inside the fused layers only
        if max(M, N, K) >= 4336 and min(M, N, K) < 4336 and (N >= 4336 or K >= 4336 or fl == 5): s_in, s_out = 4, 4   # vocabulary projection: TF32
        if fl == 2: s_out = 4                                        # residual-sum outputs (forward z, backward dx) are fp32
    aux = M * N * s_in if fl in (2, 4) else 0
    gb += n * ((M * K + N * K) * s_in + M * N * s_out + aux)
print(f"GEMM class: algorithmic bytes per step {gb / 1e9:.3f} GB over {sum(n for (c, w, t), (n, ms) in groups.items() if names[c] == 'gemm_tf32')} launches")
