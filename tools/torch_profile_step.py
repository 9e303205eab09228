"""torch.profiler view of one training step of the headline model: every CUDA kernel (ours and torch's glue) by total
device time.  python tools/torch_profile_step.py [rows]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
from speech_tranformer_pytorch_b200 import data as sdata, model as smodel, parallel as spar
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0); V = 4337
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 45
ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
tot = sum(e.device_time_total for e in ev)
print(f"total device time {tot / 1e3:.2f} ms over {sum(e.count for e in ev)} kernels")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:rows]:
    print(f"{e.device_time_total / 1e3:8.3f} ms  x{e.count:4d}  {e.key[:120]}")
