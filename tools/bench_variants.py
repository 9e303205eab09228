"""Training-step time of the other BASELINE.json training shapes on 1 x B200 (not the headline bench line):
  --layers 12                      configs[3] depth (12+12 layers), fixed T = 1000
  --ragged --frames 2000           configs[2] batch structure (lengths U[200, 2000], padded to T_max = 2000), fp32 / TF32
  --ctc                            adds the CTC head on the encoder output (joint loss, weight 0.3)
python tools/bench_variants.py [--layers 6] [--frames 1000] [--ragged] [--ctc] [--steps 5]
Under torchrun (configs[3] is quoted data-parallel) every rank takes its own synthetic shard, the gradient exchange is the
trainer's bucketed NCCL all-reduce and the time is the maximum over ranks:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_variants.py --layers 12 --ctc"""
import argparse, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
from speech_tranformer_pytorch_b200 import data as sdata, model as smodel, parallel as spar
ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=6); ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--frames", type=int, default=1000); ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--ragged", action="store_true"); ap.add_argument("--ctc", action="store_true")
a = ap.parse_args()
import torch.distributed as dist
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local); V = 4337; F = stb.functional
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(2018)
net = smodel.Transformer(smodel.headline_config(num_enc_layer=a.layers, num_dec_layer=a.layers))
smodel.init_parameters(net); net = net.to(dev).train()
ctc_proj = torch.nn.Linear(512, V).to(dev) if a.ctc else None
att = stb.LabelSmoothingLoss(0.1, V, weight=torch.ones(V, device=dev), ignore_index=0).to(dev)
crit = stb.JointCTCAttentionLoss(att, ctc_weight=0.3, blank=0) if a.ctc else att
mods = torch.nn.ModuleList([net] + ([ctc_proj] if a.ctc else []))
tr = spar.DataParallelTrainer(mods, d_model=512)
tr.broadcast_parameters(0)
inputs, targets, il, tl, truth = [t.to(dev) for t in sdata.synthetic_batch(a.batch, a.frames, 50, 80, V, seed=2018 + rank,
                                                                            fixed_len=not a.ragged, t_min=200)]
def loss_fn():
    enc, _ = net.encoder(inputs, il)
    dec, _, _ = net.decoder(targets, tl, il, enc)
    logits = F.linear(dec, net.tgt_word_proj.weight)
    if not a.ctc:
        return crit(logits.view(-1, V), truth.view(-1))
    labels = torch.where(truth > 3, truth, torch.full_like(truth, 4))
    return crit(logits.view(-1, V), truth.view(-1), F.linear(enc, ctc_proj.weight, ctc_proj.bias), labels, il, tl - 1)
for _ in range(3): tr.train_step(loss_fn)
if world > 1: dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps): loss = tr.train_step(loss_fn)
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / a.steps, float(il.sum())], device=dev, dtype=torch.float64)
if world > 1:
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    t[0] = tmax[0]
ms, valid = float(t[0]), float(t[1])
if rank == 0:
    print(json.dumps({"n_gpus": world, "layers": a.layers, "frames_max": a.frames, "ragged": a.ragged, "ctc": a.ctc, "ms_per_step": ms,
                      "padded_frames_per_s": world * a.batch * a.frames / (ms * 1e-3), "valid_frames_per_s": valid / (ms * 1e-3),
                      "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                      "buckets_started_under_backward_per_step": tr.early_launches / (a.steps + 3) if world > 1 else None}))
if world > 1:
    dist.destroy_process_group()
