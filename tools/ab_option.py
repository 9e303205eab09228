"""In-process A/B of a library option on the headline training step: alternates the values several times (same box, same
clocks, same allocator state) and prints the per-value mean / min step time.
python tools/ab_option.py <option> <value> <value> [...] [--reps 4] [--steps 5]"""
import argparse, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
from speech_tranformer_pytorch_b200 import data as sdata, model as smodel, parallel as spar
ap = argparse.ArgumentParser()
ap.add_argument("option"); ap.add_argument("values", type=int, nargs="+")
ap.add_argument("--reps", type=int, default=4); ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0); V = 4337
lib = stb._lib.load()
torch.manual_seed(2018)
net = smodel.Transformer(smodel.headline_config(dropout=0.1)); smodel.init_parameters(net); net = net.to(dev).train()
crit = stb.LabelSmoothingLoss(0.1, V, weight=torch.ones(V, device=dev), ignore_index=0).to(dev)
tr = spar.DataParallelTrainer(net, d_model=512)
batch = [t.to(dev) for t in sdata.synthetic_batch(32, 1000, 50, 80, V)]
def step():
    inputs, targets, il, tl, truth = batch
    return tr.train_step(lambda: crit(net(inputs, il, targets, tl)[0].view(-1, V), truth.view(-1)))
for _ in range(3): step()
res = {v: [] for v in a.values}
for rep in range(a.reps):
    for v in a.values:
        stb._lib.check(lib.st_set_option(a.option.encode(), v))
        step(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps): step()
        e1.record(); torch.cuda.synchronize()
        res[v].append(e0.elapsed_time(e1) / a.steps)
for v, ts in res.items():
    print(f"{a.option}={v}: mean {sum(ts) / len(ts):.3f} ms  min {min(ts):.3f}  all {[round(t, 3) for t in ts]}")
