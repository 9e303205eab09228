"""Two (or more) ranks: the C-ABI collective (st_allreduce_*, include/st_b200.h) against torch.distributed, and one training
step of the data-parallel trainer with collective="library" against collective="torch".
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_allreduce_abi.py"""
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import speech_tranformer_pytorch_b200 as stb
from speech_tranformer_pytorch_b200 import parallel as spar, model as smodel, data as sdata
if local == 0: stb.build()
dist.barrier()
coll = spar.LibraryCollective(dev)
torch.manual_seed(100 + rank)
x = torch.randn(1 << 22, device=dev); y = x.clone()
coll.all_reduce(x, async_op=False); dist.all_reduce(y)
err = (x - y).abs().max().item()
b = torch.full((1000,), float(rank), device=dev); coll.broadcast(b, 0); torch.cuda.synchronize()
ok_b = bool((b == 0).all())
# one training step, both collectives, same seeds -> identical parameters afterwards
def run(collective):
    torch.manual_seed(7)
    cfg = smodel.headline_config(num_enc_layer=1, num_dec_layer=1, d_model=128, n_heads=2, d_inner_hid=256, vocab_size=50)
    net = smodel.Transformer(cfg); smodel.init_parameters(net); net = net.to(dev).train()
    crit = stb.LabelSmoothingLoss(0.1, 50, weight=torch.ones(50, device=dev), ignore_index=0).to(dev)
    tr = spar.DataParallelTrainer(net, d_model=128, collective=collective); tr.broadcast_parameters(0)
    inputs, targets, il, tl, truth = [t.to(dev) for t in sdata.synthetic_batch(4, 120, 12, 80, 50, seed=3 + rank)]
    for layer in net.modules():
        if isinstance(layer, torch.nn.Dropout): layer.p = 0.0
    net.eval()      # no dropout: the two runs must agree bit for bit
    for p in net.parameters(): p.requires_grad_(True)
    tr.train_step(lambda: crit(net(inputs, il, targets, tl)[0].view(-1, 50), truth.view(-1)))
    torch.cuda.synchronize()
    return tr.fp.flat.clone(), tr
pa, _ = run("torch"); pb, trb = run("library")
diff = (pa - pb).abs().max().item()
coll.close(); trb.lib_collective.close()
if rank == 0:
    print(json.dumps({"world": world, "allreduce_max_abs_diff_vs_torch": err, "broadcast_ok": ok_b, "train_step_param_max_abs_diff": diff}))
dist.destroy_process_group()
