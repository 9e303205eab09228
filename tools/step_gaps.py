"""Device-timeline gaps of one training step of the headline model (torch.profiler / CUPTI): how much of the step the GPU
idles between kernels, whether the host is the limiter there (launch -> start latency), and which kernels follow the
long gaps.  python tools/step_gaps.py [option=value ...]   (options are st_set_option names, e.g. pdl=0)"""
import os, sys, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
from speech_tranformer_pytorch_b200 import data as sdata, model as smodel, parallel as spar
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0); V = 4337
lib = stb._lib.load()
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    stb._lib.check(lib.st_set_option(k.encode(), int(v)))
torch.manual_seed(2018)
net = smodel.Transformer(smodel.headline_config(dropout=0.1)); smodel.init_parameters(net); net = net.to(dev).train()
crit = stb.LabelSmoothingLoss(0.1, V, weight=torch.ones(V, device=dev), ignore_index=0).to(dev)
tr = spar.DataParallelTrainer(net, d_model=512)
batch = [t.to(dev) for t in sdata.synthetic_batch(32, 1000, 50, 80, V)]
def step():
    inputs, targets, il, tl, truth = batch
    return tr.train_step(lambda: crit(net(inputs, il, targets, tl)[0].view(-1, V), truth.view(-1)))
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
print(f"unprofiled step {e0.elapsed_time(e1) / 5:.3f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); step(); torch.cuda.synchronize()
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
             if e.device_type.name == "CUDA" and e.time_range.end > e.time_range.start), key=lambda t: t[0])
# second step only: starts at the second zero_grad memset; simply take the last half of the kernels
ks = ks[len(ks) // 2:]
span = ks[-1][1] - ks[0][0]
busy = sum(e - s for s, e, _ in ks)
gaps = [(ks[i + 1][0] - max(k[1] for k in ks[max(0, i - 3):i + 1]), ks[i][2], ks[i + 1][2]) for i in range(len(ks) - 1)]
pos = [g for g in gaps if g[0] > 0]
print(f"kernels {len(ks)}  span {span / 1e3:.3f} ms  busy {busy / 1e3:.3f} ms  idle {sum(g[0] for g in pos) / 1e3:.3f} ms")
hist = collections.Counter()
for g, _, _ in pos:
    hist[min(int(g), 20)] += 1
print("gap histogram (us: count):", " ".join(f"{k}:{hist[k]}" for k in sorted(hist)))
by_next = collections.defaultdict(lambda: [0, 0.0])
for g, prev, nxt in pos:
    by_next[nxt[:60]][0] += 1
    by_next[nxt[:60]][1] += g
print("idle time by FOLLOWING kernel:")
for name, (n, t) in sorted(by_next.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {t:8.1f} us  x{n:4d}  avg {t / n:5.2f}  {name}")
big = sorted(pos, key=lambda g: -g[0])[:12]
print("longest gaps:")
for g, prev, nxt in big:
    print(f"  {g:7.1f} us  after {prev[:50]}  before {nxt[:50]}")
