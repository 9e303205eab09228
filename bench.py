#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: frames/sec of one training step
(forward + label-smoothed CE + backward + gradient all-reduce + clip + Noam-Adam, i.e. the body of
train.py:37-46) of the 6+6-layer Speech-Transformer of BASELINE.json configs[1]
(d_model 512, 8 heads, d_ff 2048, fp32 storage / TF32 tensor cores, synthetic 80-dim fbank, B=32 per GPU, T=1000).

    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU PyTorch path on the host cores

Prints ONE JSON line (rank 0).  `value` = all ranks' frames / max-over-ranks device time with inputs resident
in HBM; `e2e` = same through the public module API with pinned-host inputs copied H2D and the loss read back
D2H every step; `roofline` = the dominant kernel (the tcgen05 TF32 GEMM) timed live with CUDA events;
`cpu_baseline` = the reference's CPU path on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec fwd+bwd at (B=32,T=1000,d=512)"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU")
    ap.add_argument("--frames", type=int, default=1000, help="T_max")
    ap.add_argument("--targets", type=int, default=50, help="L_max")
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--cpu-sample-batch", type=int, default=8, help="utterances per CPU-baseline step (BASELINE.md §5: B=8)")
    ap.add_argument("--dtype", default=os.environ.get("ST_BENCH_DTYPE", "fp32"), choices=["fp32", "tf32", "fp16", "bf16"],
                    help="fp32 = fp32 model (BASELINE.json configs[1]): fp32 tensors at every module boundary, fp16 tensor-core "
                         "operands inside MultiHeadAttention / PositionwiseFeedForward (functional.set_fp32_engine('fp16'), the "
                         "library default); tf32 = the same fp32 model with TF32 operands throughout; fp16 / bf16 = 16-bit "
                         "activations end to end, fp32 accumulate (configs[2])")
    ap.add_argument("--ragged", action="store_true", help="utterance lengths U[200, frames] padded to --frames (configs[2])")
    ap.add_argument("--no-variants", action="store_true",
                    help="skip the extra measurements of the default run: the same step with fp16 / bf16 operands, BASELINE.json "
                         "configs[2] (bf16, ragged T <= 2000), and the reference's modules in eager PyTorch on this GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU reference
def cpu_reference_step_fn(args, batch):
    """The reference's own training step on CPU: oracle/_ref (patched scratch copy of the reference package,
    kind 'reference') when present, else the oracle port (kind 'port').  Returns (step_fn, kind, frames/step)."""
    import torch
    import types
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    from oracle import st_oracle as O
    inputs, targets, in_len, tgt_len, truth = O.synthetic_batch(batch, args.frames, args.targets, 80, 4337, seed=2018)
    V, d = 4337, 512
    cfgd = dict(feature_dim=80, vocab_size=V, max_inputs_length=2048, max_target_length=64, d_model=d, n_heads=8, d_k=64,
                d_v=64, d_inner_hid=2048, num_enc_layer=args.layers, num_dec_layer=args.layers, dropout=args.dropout,
                emb_scale=1, return_attns=False)
    if os.path.isdir(os.path.join(ref_dir, "transformer")) and not os.environ.get("ST_BENCH_FORCE_PORT"):
        kind = "reference"
        for name in ("editdistance", "matplotlib", "matplotlib.pyplot", "tensorboardX"):
            sys.modules.setdefault(name, types.ModuleType(name))
        for k in [k for k in sys.modules if k == "transformer" or k.startswith("transformer.")]:
            del sys.modules[k]
        sys.path.insert(0, ref_dir)
        from transformer.Models import Transformer
        from transformer.Loss import LabelSmoothingLoss
        from transformer.Optim import ScheduledOptim
        from transformer.Utils import AttrDict, init_parameters
        sys.path.pop(0)
        torch.manual_seed(2018)
        model = Transformer(AttrDict(cfgd))
        init_parameters(model)
        model.train()
        crit = LabelSmoothingLoss(0.1, V, weight=torch.ones(V), size_average=True, ignore_index=0)
        optim = ScheduledOptim(model, d, AttrDict({"n_warmup_steps": 12000}))
        state = {"step": 0}

        def step():
            state["step"] += 1
            optim.zero_grad()
            logits, _ = model(inputs, in_len, targets, tgt_len)                                  # train.py:39
            loss = crit(logits.contiguous().view(-1, V), truth.contiguous().view(-1))            # train.py:40
            loss.backward()                                                                      # train.py:44
            torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)                              # train.py:45
            optim.step(state["step"])                                                            # train.py:46
            return float(loss)
    else:
        kind = "port"
        from oracle import model_port
        step = model_port.make_train_step(cfgd, inputs, targets, in_len, tgt_len, truth)
    return step, kind, batch * args.frames


def run_cpu(args, steps, warmup, batch):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind, frames = cpu_reference_step_fn(args, batch)
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": frames * steps / total, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{batch} utterances x T={args.frames} (of B={args.batch}) per step, full {args.layers}+{args.layers} model, "
                      f"train mode dropout {args.dropout}, fwd+loss+bwd+clip+Adam, {steps} timed steps after {warmup} warm-up",
            "ms_per_step": 1e3 * total / steps}


DTYPE_DESC = {"fp32": "f32 (fp32 parameters, activations and gradients at every module boundary; inside MultiHeadAttention / "
                      "PositionwiseFeedForward the tensor-core operands are fp16 = TF32's 10-bit mantissa, fp32 accumulate / "
                      "statistics, per-operator power-of-two gradient scale derived on the device; frontend, vocabulary "
                      "projection: TF32 operands; loss, optimizer: fp32)",
              "tf32": "tf32 (fp32 storage, TF32 tensor-core operands, fp32 accumulate)",
              "fp16": "fp16 (fp16 activations and tensor-core operands = TF32's 10-bit mantissa, fp32 accumulate / statistics / "
                      "parameters / optimizer, loss scale 2^14)",
              "bf16": "bf16 (bf16 activations and tensor-core operands, fp32 accumulate / statistics / parameters / optimizer)"}


# what the parity tests establish for each --dtype (tolerances are written in the tests; measured values: DESIGN.md section 2,
# gpurun_out/round2_parity_errors.json / half_parity_errors.json written by the GPU tests)
PRECISION_NOTE = {
    "fp32": "fp32 tensors in / out of every module; fp16 operands carry TF32's 11 significant bits (tcgen05 has no fp32 MMA). "
            "Every fp32-tolerance parity test (1e-3 per module, 3e-3 for the composed 6+6 x 512 model) runs under this engine AND "
            "under the TF32 engine (tests/conftest.py `engine`); measured on the 6+6 x 512 model vs the float64 oracle: logits "
            "1.07e-3, loss 8e-6, parameter gradients 2.1e-3 (TF32 engine: 1.24e-3, 5e-5, 2.2e-3)",
    "tf32": "fp32 tensors, TF32 operands (round 1's arithmetic): 1e-3 per module, 3e-3 composed; measured 1.24e-3 / 5e-5 / 2.2e-3",
    "fp16": "fp16 activations end to end, loss scale 2^14: 1.5e-3 per module, 3e-3 composed; measured 1.0e-3 / 7.7e-5 / 2.5e-3",
    "bf16": "bf16 activations end to end: 2e-2 per module, 3e-2 composed (8-bit mantissa); measured 6.6e-3 / 4e-4 / 7.0e-3",
}


def workload_config(args, n):
    which = "configs[2]" if (args.ragged or args.dtype == "bf16") else "configs[1]"
    lens = f"T in U[200,{args.frames}] padded to {args.frames}" if args.ragged else f"T={args.frames}"
    return {"workload": f"BASELINE.json {which}: {args.layers}+{args.layers}-layer enc/dec d_model=512 h=8 d_ff=2048, "
                        f"{'fp32 tensors / fp16 operands in the fused layers' if args.dtype == 'fp32' else args.dtype + ' operands'}, synthetic 80-dim fbank B={args.batch}/GPU {lens} L<={args.targets} V=4337",
            "step": "fwd + label-smoothed CE + bwd + grad all-reduce + clip + Noam-Adam (train.py:37-46)",
            "dropout": args.dropout, "per_gpu_batch": args.batch, "global_batch": args.batch * n, "parallelism": f"dp{n}",
            "l2": "per-step working set (several GB of activations) >> 126 MB L2, no explicit flush needed"}


# ------------------------------------------------------------------------------------------------ reference arm
def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.gpus
    res = run_cpu(args, max(1, args.steps), max(0, args.warmup), args.cpu_sample_batch)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, n),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0



# ------------------------------------------------------------------------------------------------ variants
def measure_variant(stb, smodel, spar, sdata, dev, args, dtype, ragged, frames, steps=5, warmup=3):
    """ms/step and frames/s of the same training step with other operand types / batch shapes (rank-0, single GPU), plus the
    EncoderLayer forward + backward alone.  Used for the `variants` block of the default run's JSON line."""
    import torch
    V, d = 4337, 512
    stb.functional.set_fp32_engine("tf32" if dtype == "tf32" else "fp16")     # main() restores its own setting afterwards
    cfg = smodel.headline_config(num_enc_layer=args.layers, num_dec_layer=args.layers, dropout=args.dropout, compute_dtype=dtype,
                                 max_inputs_length=max(2048, frames))
    act = smodel.COMPUTE_DTYPES[dtype]
    torch.manual_seed(2018)
    net = smodel.Transformer(cfg)
    smodel.init_parameters(net)
    net = net.to(dev).train()
    crit = stb.LabelSmoothingLoss(0.1, V, weight=torch.ones(V, device=dev), size_average=True, ignore_index=0).to(dev)
    trainer = spar.DataParallelTrainer(net, d_model=d, n_warmup_steps=12000, max_grad_norm=5.0, compute_dtype=act, overlap=False)
    host = sdata.synthetic_batch(args.batch, frames, args.targets, 80, V, seed=2018, fixed_len=not ragged, t_min=min(200, frames))
    batch = [t.to(dev) for t in host]

    def step():
        inputs, targets, in_len, tgt_len, truth = batch
        return trainer.train_step(lambda: crit(net(inputs, in_len, targets, tgt_len)[0].view(-1, V), truth.view(-1)))

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for _ in range(warmup):
        loss = step()
    ms = timed(step, steps)
    layer = net.encoder.layer_stack[0]
    lx = torch.randn(args.batch, frames, d, device=dev).to(act)
    lg = torch.randn(args.batch, frames, d, device=dev).to(act)
    lmask = stb.functional.LengthMask(batch[2], frames, frames)
    lparams = list(layer.parameters())

    def layer_step():
        for q in lparams:
            q.grad = None
        xin = lx.detach().requires_grad_()
        y, _ = layer(xin, slf_attn_mask=lmask)
        y.backward(lg)

    for _ in range(3):
        layer_step()
    lms = timed(layer_step, 10)
    n_tok = args.batch * frames
    flops = 3.0 * (8.0 * n_tok * d * d + 4.0 * args.batch * 8 * frames * frames * 64 + 4.0 * n_tok * d * 2048)
    out = {"dtype": dtype, "ragged": bool(ragged), "frames": frames, "ms_per_step": ms, "frames_per_s": n_tok / (ms * 1e-3),
           "valid_frames_per_s": int(host[2].sum()) / (ms * 1e-3), "loss_finite": bool(torch.isfinite(loss.detach()).item()),
           "encoder_layer_ms_fwd_bwd": lms, "encoder_layer_tflops": flops / (lms * 1e-3) / 1e12}
    del net, trainer, batch, lx, lg
    torch.cuda.empty_cache()
    return out


def measure_eager_reference_on_gpu(dev, args, steps=3, warmup=2):
    """SURVEY.md §8(d) 'recommended extra': the reference's OWN modules (oracle/_ref, unmodified hot-path files) in eager PyTorch
    on this GPU (cuBLAS / cuDNN library kernels, fp32 with TF32 matmuls allowed) — shows the gain over the library path."""
    import torch
    import types
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "transformer")):
        return {"unavailable": "oracle/_ref not present"}
    for name in ("editdistance", "matplotlib", "matplotlib.pyplot", "tensorboardX"):
        sys.modules.setdefault(name, types.ModuleType(name))
    saved = {k: v for k, v in sys.modules.items() if k == "transformer" or k.startswith("transformer.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, ref_dir)
    try:
        from transformer.Models import Transformer
        from transformer.Utils import AttrDict, init_parameters
        import transformer.Utils as U
        from oracle import st_oracle as O
        V, d = 4337, 512
        cfgd = dict(feature_dim=80, vocab_size=V, max_inputs_length=2048, max_target_length=64, d_model=d, n_heads=8, d_k=64,
                    d_v=64, d_inner_hid=2048, num_enc_layer=args.layers, num_dec_layer=args.layers, dropout=args.dropout,
                    emb_scale=1, return_attns=False)
        torch.manual_seed(2018)
        model = Transformer(AttrDict(cfgd))
        init_parameters(model)
        model = model.to(dev).train()
        # Utils.py:41-70 builds its masks with numpy on the host and returns CPU tensors: move them where the model lives
        pm, fm = U.padding_info_mask, U.feature_info_mask
        import transformer.Models as M
        M.padding_info_mask = lambda a, b: pm(a.cpu(), b.cpu()).to(dev)
        M.feature_info_mask = lambda a: fm(a.cpu()).to(dev)
        crit = torch.nn.CrossEntropyLoss(ignore_index=0)                 # train.py:120 (Loss.py's class is CPU-only as written)
        opt = torch.optim.Adam(model.parameters(), betas=(0.9, 0.98), eps=1e-9)
        inputs, targets, in_len, tgt_len, truth = [t.to(dev) for t in O.synthetic_batch(args.batch, args.frames, args.targets, 80, V, seed=2018)]
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True

        def step():
            opt.zero_grad()
            logits, _ = model(inputs, in_len, targets, tgt_len)
            loss = crit(logits.reshape(-1, V), truth.reshape(-1))
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
            opt.step()

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        torch.backends.cuda.matmul.allow_tf32 = old
        ms = e0.elapsed_time(e1) / steps
        peak = torch.cuda.max_memory_allocated(dev) / 2 ** 30
        del model, opt
        torch.cuda.empty_cache()
        return {"ms_per_step": ms, "frames_per_s": args.batch * args.frames / (ms * 1e-3), "peak_gib": peak,
                "what": "the reference's Transformer (oracle/_ref: Attention.py / SubLayers.py / Layers.py / Models.py as patched by "
                        "oracle/make_ref.py) in eager PyTorch on this GPU, allow_tf32, nn.CrossEntropyLoss + clip + Adam"}
    except Exception as e:       # a baseline row must never take the bench down
        return {"unavailable": repr(e)[:300]}
    finally:
        sys.path.remove(ref_dir)
        for k in [k for k in sys.modules if k == "transformer" or k.startswith("transformer.")]:
            del sys.modules[k]
        sys.modules.update(saved)


# ------------------------------------------------------------------------------------------------ B200 arm
def main_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = world

    import speech_tranformer_pytorch_b200 as stb
    from speech_tranformer_pytorch_b200 import data as sdata, model as smodel, parallel as spar
    if local == 0:
        stb.build()                      # one builder per node: ranks must not run nvcc into the same objects concurrently
    if world > 1:
        dist.barrier()
    lib = stb._lib.load()
    stb._lib.check(lib.st_device_check(local))
    stb.functional.set_fp32_engine("tf32" if args.dtype == "tf32" else "fp16")

    V, d = 4337, 512
    cfg = smodel.headline_config(num_enc_layer=args.layers, num_dec_layer=args.layers, dropout=args.dropout,
                                 compute_dtype=args.dtype)
    act_dtype = smodel.COMPUTE_DTYPES[args.dtype]
    torch.manual_seed(2018)
    net = smodel.Transformer(cfg)
    smodel.init_parameters(net)
    net = net.to(dev).train()
    crit = stb.LabelSmoothingLoss(0.1, V, weight=torch.ones(V, device=dev), size_average=True, ignore_index=0).to(dev)
    trainer = spar.DataParallelTrainer(net, d_model=d, n_warmup_steps=12000, max_grad_norm=5.0, compute_dtype=act_dtype)
    trainer.broadcast_parameters(0)

    host = sdata.synthetic_batch(args.batch, args.frames, args.targets, 80, V, seed=2018 + rank, pin=True,
                                 fixed_len=not args.ragged, t_min=min(200, args.frames))
    resident = [t.to(dev) for t in host]

    def step_on(inputs, targets, in_len, tgt_len, truth):
        def loss_fn():
            logits, _ = net(inputs, in_len, targets, tgt_len)
            return crit(logits.view(-1, V), truth.view(-1))
        return trainer.train_step(loss_fn)

    def step_resident():
        return step_on(*resident)

    h2d = sum(t.numel() * t.element_size() for t in host)

    prefetch = sdata.DevicePrefetcher(dev)

    def step_e2e():
        # every step copies ITS inputs from pinned host memory (submitted one step ahead on a side stream, so the PCIe
        # transfer runs under the previous step, as a pin_memory DataLoader does) and reads its loss back to the host
        batch = prefetch.get()
        loss = step_on(*batch)                                    # enqueue the step first: the GPU is idle after the last .item()
        prefetch.submit(host)                                     # next step's inputs, copied under this step
        return loss.item()                                        # device -> host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        barrier()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = lib.st_launch_count()
    early0 = trainer.early_launches
    ms = timed(step_resident, args.steps)
    peak_mem_gib = torch.cuda.max_memory_allocated(dev) / 2 ** 30      # model + optimizer state + one step's activations
    launches = lib.st_launch_count() - launches0
    early = (trainer.early_launches - early0) / args.steps
    clocks = sampler.stop() if sampler else None
    frames = args.batch * args.frames * n * args.steps
    value = frames / (ms * 1e-3)
    valid_frames = int(host[2].sum().item())      # this rank's un-padded frames per step (== batch * frames unless --ragged)

    prefetch.submit(host)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    prefetch.get()                                                # drain the one batch in flight
    e2e = {"value": frames / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
           "ms_per_step": ms_e2e / args.steps,
           "note": "public API (model + loss modules + DataParallelTrainer.train_step) on a pinned host batch: H2D copy of every "
                   "step's inputs on a side stream one step ahead (data.DevicePrefetcher), loss.item() every step"}

    # ---- N > 1: the gradient exchange (bucketed all-reduce started under backward, parallel.py).  One untimed step checks
    # that every rank holds the SAME reduced gradient: a bucket reduced before its last local write would differ.
    dp = None
    if world > 1:
        trainer.zero_grad()
        inputs, targets, in_len, tgt_len, truth = resident
        logits, _ = net(inputs, in_len, targets, tgt_len)
        crit(logits.view(-1, V), truth.view(-1)).backward()
        trainer.allreduce_gradients()
        ref = trainer.fp.grad.clone()
        dist.broadcast(ref, 0)
        diff = (ref - trainer.fp.grad).abs().max()
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        dp = {"collective": "NCCL all-reduce(SUM) of the flat fp32 gradient buffer", "bytes_per_step": 4 * trainer.fp.numel,
              "buckets": len(trainer.buckets.items), "buckets_started_under_backward_per_step": early,
              "replica_grad_max_abs_diff": float(diff.item())}
        if dp["replica_grad_max_abs_diff"] != 0.0:
            print(f"bench.py: WARNING replicas disagree on the reduced gradient ({dp['replica_grad_max_abs_diff']:.3e})",
                  file=sys.stderr, flush=True)

    # ---- roofline of the dominant kernel, timed live with CUDA events on the launching stream
    roofline, breakdown = None, None
    if not args.no_roofline:
        import ctypes as C
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else \
            "fallback 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md)"
        # TF32 dense peak is not in MEASURED_PEAKS.json: measure it the same way (8192^3 torch.matmul, allow_tf32)
        a = torch.randn(8192, 8192, device=dev)
        b = torch.randn(8192, 8192, device=dev)
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        torch.backends.cuda.matmul.allow_tf32 = old
        tf32_peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
        del a, b
        # per-class times are CUDA-event intervals around every launch: measured with the library's side streams off, so no
        # interval contains another kernel's work (the timed region above runs with them on)
        side_on = os.environ.get("ST_SIDE_STREAMS", "1") != "0"
        lib.st_set_option(b"side_streams", 0)
        lib.st_profile_reset()
        lib.st_profile_enable(1)
        psteps = min(args.steps, 3)
        ms_prof = timed(step_resident, psteps)
        lib.st_profile_enable(0)
        lib.st_set_option(b"side_streams", 1 if side_on else 0)
        breakdown = {}
        for c in range(lib.st_profile_classes()):
            t, w, k = C.c_double(), C.c_double(), C.c_int64()
            stb._lib.check(lib.st_profile_read(c, C.byref(t), C.byref(w), C.byref(k)))
            if k.value:
                breakdown[lib.st_profile_class_name(c).decode()] = {
                    "launches_per_step": k.value / psteps, "ms_per_step": t.value / psteps,
                    "share_of_step": t.value / ms_prof, "work_per_step": w.value / psteps,
                    "rate": w.value / (t.value * 1e-3) / 1e12 if t.value > 0 else None}
        lib.st_profile_reset()
        g = breakdown.get("gemm_tf32")
        traffic, traffic_note = None, None
        alg_bytes = None
        try:   # DRAM bytes per launch of the same kernels from the committed ncu launch list of one step (profiles/), averaged
               # over EVERY GEMM launch of the step, next to the algorithmic bytes (operands + output + aux operand, once each)
            with open(os.path.join(ROOT, "profiles", f"r2_gemm_traffic_{args.dtype if args.dtype in ('tf32', 'fp32') else 'bf16'}.json")) as f:
                tj = json.load(f)
            traffic, traffic_note, alg_bytes = tj["dram_bytes_per_launch"], tj["source"], tj.get("algorithmic_bytes_per_launch")
        except Exception:
            pass
        if g:
            roofline = {"kernel": f"gemm_kernel / gemm_2sm_kernel (tcgen05 kind::{'tf32' if args.dtype == 'tf32' else 'f16'}, all projection / FFN / gradient GEMMs)",
                        "bound": "tensor", "achieved": g["rate"], "peak": peak, "unit": "TFLOP/s",
                        "frac": g["rate"] / peak, "traffic": traffic, "traffic_algorithmic": alg_bytes,
                        "traffic_source": traffic_note, "peak_source": peak_src,
                        "peak_tf32_measured": tf32_peak, "frac_of_tf32_peak": g["rate"] / tf32_peak,
                        "flops_per_launch": g["work_per_step"] / g["launches_per_step"],
                        "avg_launch_ms": g["ms_per_step"] / g["launches_per_step"],
                        "share_of_step": g["share_of_step"],
                        "note": "achieved = algorithmic 2MNK FLOPs / CUDA-event time over every GEMM launch of the profiled "
                                "steps; TF32 operands run at half the bf16 MMA rate (frac against the bf16 peak is capped at 0.5 "
                                "for --dtype tf32)"}

    # ---- the fused EncoderLayer alone (north_star: tensor-pipe utilisation of EncoderLayer fwd+bwd at B=32,T=1000,d=512,h=8)
    enc_layer = None
    if rank == 0 and not args.no_roofline:
        # a private copy of the first encoder layer with its own flat buffers (as in training: operand-precision weight twins,
        # parameter gradients written straight into the flat gradient buffer) — at N > 1 only rank 0 runs this leg, so it
        # must not touch the data-parallel trainer's buckets
        import copy
        layer = copy.deepcopy(net.encoder.layer_stack[0])
        ltrainer = spar.DataParallelTrainer(layer, d_model=d, compute_dtype=act_dtype, overlap=False)
        lx = torch.randn(args.batch, args.frames, d, device=dev).to(act_dtype)
        lg = torch.randn(args.batch, args.frames, d, device=dev).to(act_dtype)
        lmask = stb.functional.LengthMask(resident[2], args.frames, args.frames)

        def layer_step():
            ltrainer.zero_grad()
            xin = lx.detach().requires_grad_()
            y, _ = layer(xin, slf_attn_mask=lmask)
            y.backward(lg)

        for _ in range(3):
            layer_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            layer_step()
        e1.record()
        torch.cuda.synchronize()
        lms = e0.elapsed_time(e1) / reps
        N_tok, T_, h_, dk_, dff_ = args.batch * args.frames, args.frames, 8, 64, 2048
        flops = 3.0 * (8.0 * N_tok * d * d + 4.0 * args.batch * h_ * T_ * T_ * dk_ + 4.0 * N_tok * d * dff_)
        tf = flops / (lms * 1e-3) / 1e12
        enc_layer = {"ms_fwd_bwd": lms, "algorithmic_gflop": flops / 1e9, "tflops": tf,
                     "frames_per_s": N_tok / (lms * 1e-3),
                     "frac_of_tf32_peak_measured": tf / tf32_peak if tf32_peak else None,
                     "frac_of_bf16_peak": tf / peak,
                     "frac_of_bf16_peak_burst": tf / float(peaks.get("bf16_tflops", 1660.0)) if peaks else None,
                     "dtype": args.dtype,
                     "note": "one EncoderLayer (MHA + FFN, train mode, dropout 0.1) forward + backward, B x T x 512 resident in HBM; "
                             "FLOPs = 3 x (8Nd^2 + 4BhT^2dk + 4Nd*dff), no recompute counted, padded frames counted as the "
                             "reference computes them; --dtype fp32: fp32 in / out, fp16 operands inside; tf32 operands issue at half the "
                             "bf16 / fp16 tensor-pipe rate"}
        del ltrainer, layer
        trainer.zero_grad()

    cpu = None
    if rank == 0 and n == 1 and not args.no_cpu_baseline:
        try:
            cpu = run_cpu(args, steps=2, warmup=1, batch=args.cpu_sample_batch)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the GPU numbers stand on their own
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)[:200]}

    variants, eager = None, None
    if rank == 0 and n == 1 and not args.no_variants and not args.no_roofline:
        del resident
        torch.cuda.empty_cache()
        variants = []
        for dt_, ragged_, frames_ in (("fp32", False, args.frames), ("tf32", False, args.frames), ("fp16", False, args.frames),
                                      ("bf16", False, args.frames), ("bf16", True, 2 * args.frames)):
            if (dt_, ragged_, frames_) == (args.dtype, args.ragged, args.frames):
                continue
            try:
                variants.append(measure_variant(stb, smodel, spar, sdata, dev, args, dt_, ragged_, frames_))
            except Exception as e:
                variants.append({"dtype": dt_, "ragged": ragged_, "frames": frames_, "error": repr(e)[:200]})
        stb.functional.set_fp32_engine("tf32" if args.dtype == "tf32" else "fp16")
        eager = measure_eager_reference_on_gpu(dev, args)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": DTYPE_DESC[args.dtype], "data": "synthetic", "config": workload_config(args, n),
                "valid_frames_per_s": valid_frames * n * args.steps / (ms * 1e-3),
                "peak_mem_gib": peak_mem_gib,
                "precision": PRECISION_NOTE.get(args.dtype),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "kernel_breakdown": breakdown, "encoder_layer": enc_layer, "cpu_baseline": cpu, "data_parallel": dp,
                "variants": variants, "eager_pytorch_on_gpu": eager}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(main_reference(a) if a.impl == "reference" else main_b200(a))
