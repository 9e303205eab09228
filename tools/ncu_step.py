"""One training step of the headline model bracketed by cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:<kernel> -c 2 -o gpurun_out/<name> python tools/ncu_step.py"""
import argparse, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
from speech_tranformer_pytorch_b200 import data as sdata, model as smodel, parallel as spar
ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=6); ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--frames", type=int, default=1000); ap.add_argument("--dropout", type=float, default=0.1)
ap.add_argument("--dtype", default="fp32", choices=["fp32", "tf32", "fp16", "bf16"])  # as bench.py --dtype
a = ap.parse_args()
dev = torch.device("cuda", 0); V = 4337
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
for _ in range(2): step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
