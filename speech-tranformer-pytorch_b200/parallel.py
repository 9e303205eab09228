"""Data-parallel training step around one flat parameter / gradient buffer.

Mirrors the reference's only parallelism strategy, train_multi.py: one process per GPU (:128), parameters
broadcast from rank 0 once (:176-177), every step the gradients are all-reduced (average) across ranks
(Horovod DistributedOptimizer, :161-163), then clip + Noam-Adam (train.py:45-46, Optim.py:9-45).  Here all
parameters live in ONE contiguous fp32 buffer and all gradients in another, so the gradient exchange is an
NCCL all-reduce over NVLink/NVSwitch of that buffer and the update is two HBM-bound kernels (squared norm,
fused clip+Adam) from libst_b200.so.  The buffer is reduced in a few contiguous buckets (~13 MB): a bucket's
all-reduce starts on NCCL's stream as soon as the backward operators that write its slice have been enqueued
(`GradSink.owner`, `functional._notify`), so the exchange runs under the rest of the backward pass and only the
last bucket (front-end + first encoder layer) is exposed; whatever was not started early is reduced after
backward.  The clip uses the REDUCED gradient (the reference clips local gradients before Horovod has
synchronised them, train_multi.py:66 — a bug not reproduced here).

The flattening / bucketing logic is device-agnostic and is exercised on CPU with gloo (tests/test_dp_gloo.py);
`step()` needs the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def noam_lr(d_model: int, n_warmup_steps: int, step: int) -> float:
    """Optim.update_learning_rate (Optim.py:36-45): d^-0.5 * min(step^-0.5, step * warmup^-1.5)."""
    return d_model ** -0.5 * min(step ** -0.5, step * n_warmup_steps ** -1.5)


def ordered_parameters(module: torch.nn.Module) -> List[torch.nn.Parameter]:
    """The module's parameters in the order the flat buffer stores them: registration order, except that each
    MultiHeadAttention keeps [Wq; Wk; Wv] and [bq; bk; bv] adjacent so that st_mha_bwd can produce the packed
    projection gradients with ONE weight-gradient GEMM / column sum directly inside the flat gradient buffer."""
    seen, out = set(), []

    def add(p):
        if p is not None and id(p) not in seen:
            seen.add(id(p))
            out.append(p)

    for m in module.modules():
        if all(hasattr(m, n) for n in ("linear_q", "linear_k", "linear_v", "output_linear", "layernorm")):
            for p in (m.linear_q.weight, m.linear_k.weight, m.linear_v.weight, m.linear_q.bias, m.linear_k.bias,
                      m.linear_v.bias):
                add(p)
        for p in m.parameters(recurse=False):
            add(p)
    return out


class FlatParams:
    """Re-homes a module's parameters into one flat buffer (and their .grad into another).  Every parameter gets a
    GradSink (functional.attach_grad_sink) pointing at its slice of the gradient buffer: the composite backward
    operators write parameter gradients there directly instead of going through autograd's accumulation."""

    def __init__(self, params: Iterable[torch.nn.Parameter], twin_dtype=torch.float32):
        self.twin_dtype = twin_dtype      # element type of the operand-precision twin: float32 (TF32-rounded), float16, bfloat16
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("all parameters must share one device and dtype")
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4          # keep every parameter 16-byte aligned
        self.numel = n
        self.flat = torch.zeros(n, device=dev, dtype=dt)
        self.grad = torch.zeros(n, device=dev, dtype=dt)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                self.flat[o:o + p.numel()].copy_(p.reshape(-1))
                p.data = self.flat[o:o + p.numel()].view_as(p)
                p.grad = self.grad[o:o + p.numel()].view_as(p)
        from .functional import attach_grad_sink
        self.sinks = [attach_grad_sink(p, p.grad) for p in self.params]
        # operand-precision twin of the whole parameter buffer (TF32-rounded fp32, or fp16 / bf16 for the 16-bit path): the
        # fused Adam kernel rewrites it every step, so the attention / FFN operators never round, convert or re-pack a
        # weight matrix themselves
        self.flat_tf32 = torch.empty_like(self.flat, dtype=twin_dtype) if self.flat.is_cuda else None
        self.refresh_rounded()

    def refresh_rounded(self) -> None:
        """Recompute the twins from the fp32 parameters (after a broadcast, load_state_dict, manual edits).  Parameters
        owned by a FlatParams may only be modified through the trainer's optimizer step, or by in-place PyTorch edits
        FOLLOWED by this call: writes through `.data` / raw pointers do not bump the version counter the twins are checked
        against."""
        if self.flat_tf32 is None:
            return
        from . import _lib
        from .functional import ACT_DTYPES, _stream, attach_tf32_twin
        lib = _lib.load()
        if self.twin_dtype == torch.float32:
            _lib.check(lib.st_round_tf32(self.flat.data_ptr(), self.numel, self.flat_tf32.data_ptr(), self.numel, 1, self.numel,
                                         _stream()))
        else:
            _lib.check(lib.st_cast(self.flat.data_ptr(), _lib.DTYPE_F32, self.numel, self.flat_tf32.data_ptr(),
                                   ACT_DTYPES[self.twin_dtype], self.numel, 1, self.numel, 1.0, _stream()))
        for p, o in zip(self.params, self.offsets):
            if p.dim() >= 2:
                attach_tf32_twin(p, self.flat_tf32[o:o + p.numel()].view_as(p))

    def zero_grad(self) -> None:
        self.grad.zero_()                                # one memset; direct writers overwrite, autograd accumulates
        for s in self.sinks:                             # (kept lean: in a synchronous loop the GPU idles behind this)
            s.written = False
            s.zeroed = True                              # operators skip their own per-tensor clears this step
            s.uses = s.done = 0
        for p, o, s in zip(self.params, self.offsets, self.sinks):
            if p.grad is not s.view:                     # someone replaced / dropped .grad (zero_grad(set_to_none=True), ...)
                p.grad = self.grad[o:o + p.numel()].view_as(p)
                s.view = p.grad


class _Bucket:
    __slots__ = ("lo", "hi", "sinks", "remaining", "started", "work")

    def __init__(self, lo: int, hi: int, sinks):
        self.lo, self.hi, self.sinks = lo, hi, sinks
        self.remaining, self.started, self.work = len(sinks), False, None


class GradBuckets:
    """Contiguous slices of the flat gradient buffer (whole parameters, >= `bucket_floats` each except the last) with
    the bookkeeping that tells when a slice is final during backward.  Pure host logic (tested on CPU)."""

    def __init__(self, fp: FlatParams, bucket_floats: int):
        self.items: List[_Bucket] = []
        self.numel = fp.numel
        lo, sinks = 0, []
        for i, (p, o, s) in enumerate(zip(fp.params, fp.offsets, fp.sinks)):
            sinks.append(s)
            end = fp.offsets[i + 1] if i + 1 < len(fp.params) else fp.numel
            if end - lo >= bucket_floats or i + 1 == len(fp.params):
                self.items.append(_Bucket(lo, end, sinks))
                lo, sinks = end, []
        self.of = {id(s): b for b in self.items for s in b.sinks}

    def reset(self) -> None:
        for b in self.items:
            b.remaining, b.started, b.work = len(b.sinks), False, None

    def mark_done(self, sinks) -> List[_Bucket]:
        """Record that these sinks were written directly; returns the buckets that thereby became final: every parameter
        in them was used by a forward operator and each use has been written (GradSink.done == uses >= 1)."""
        ready = []
        for s in sinks:
            b = self.of.get(id(s))
            if b is None or b.started or s.uses < 1 or s.done != s.uses:
                continue
            b.remaining -= 1
            if b.remaining == 0:
                b.started = True
                ready.append(b)
        return ready

    def pending_ranges(self):
        """Maximal contiguous [lo, hi) ranges of buckets whose reduction has not been started."""
        out = []
        for b in self.items:
            if b.started:
                continue
            if out and out[-1][1] == b.lo:
                out[-1][1] = b.hi
            else:
                out.append([b.lo, b.hi])
        return [(lo, hi) for lo, hi in out]


class LibraryCollective:
    """The gradient exchange through the C ABI (st_allreduce_*: ncclAllReduce on a communicator the library owns) instead of
    torch.distributed — what a host that is not PyTorch would call (include/st_b200.h).  torch.distributed is used ONCE, to
    hand rank 0's NCCL unique id to the other ranks (any out-of-band channel does).  Collectives run on a side stream that
    waits for the kernels enqueued so far on the compute stream, like torch's NCCL process group does."""

    def __init__(self, device, group=None):
        from . import _lib
        self.lib = _lib.load()
        self._check = _lib.check
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n = self.lib.st_allreduce_id_bytes()
        uid = (C.c_char * n)()
        if rank == 0:
            self._check(self.lib.st_allreduce_unique_id(uid))
        box = [bytes(uid.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        comm = C.c_void_p()
        with torch.cuda.device(device):
            self._check(self.lib.st_allreduce_init(box[0], world, rank, C.byref(comm)))
        self.comm, self.device = comm, device
        self.stream = torch.cuda.Stream(device)

    class _Work:
        def __init__(self, event):
            self.event = event

        def wait(self):
            torch.cuda.current_stream().wait_event(self.event)

    def all_reduce(self, view: torch.Tensor, async_op: bool):
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        self.stream.wait_event(ready)
        self._check(self.lib.st_allreduce_run(self.comm, view.data_ptr(), view.numel(), C.c_void_p(self.stream.cuda_stream)))
        done = torch.cuda.Event()
        done.record(self.stream)
        work = LibraryCollective._Work(done)
        if async_op:
            return work
        work.wait()
        return None

    def broadcast(self, view: torch.Tensor, src: int):
        self._check(self.lib.st_allreduce_broadcast(self.comm, view.data_ptr(), view.numel(), src,
                                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def close(self):
        if self.comm:
            self._check(self.lib.st_allreduce_destroy(self.comm))
            self.comm = C.c_void_p()


def _twin_dtype_for(module: torch.nn.Module, compute_dtype):
    """Element type of the parameter twins: the compute type, except that fp32 modules whose composite operators run the
    fp16-operand engine (functional.set_fp32_engine; every attention head 64 wide) keep fp16 twins."""
    from .functional import fp32_engine
    if compute_dtype != torch.float32 or fp32_engine() != "fp16":
        return compute_dtype
    head_sizes = {m.d_k for m in module.modules() if hasattr(m, "d_k") and hasattr(m, "n_head")}
    return torch.float16 if head_sizes <= {64} else torch.float32


class DataParallelTrainer:
    """zero_grad -> (caller: forward + backward) -> allreduce -> clip + Adam, on flat buffers."""

    def __init__(self, module: torch.nn.Module, d_model: int, n_warmup_steps: int = 12000, max_grad_norm: float = 5.0,
                 betas=(0.9, 0.98), eps: float = 1e-9, process_group=None, overlap: bool = True, bucket_mb: float = 13.0,
                 compute_dtype=torch.float32, loss_scale: Optional[float] = None, collective: str = "torch"):
        """compute_dtype: activation type the module runs in (float32 = TF32 operands; float16 / bfloat16 = 16-bit operands;
        the parameter twins follow it).  loss_scale: factor applied to the loss before backward and divided out inside the
        Adam kernel (a power of two; default 2**14 for float16 — whose activation gradients would otherwise underflow —
        and 1 otherwise).  A step whose gradient norm is not finite is skipped by the kernel (st_adam_step)."""
        if collective not in ("torch", "library"):
            raise ValueError("collective must be 'torch' (torch.distributed) or 'library' (st_allreduce_* of the C ABI)")
        self.module = module
        self.compute_dtype = compute_dtype
        self.loss_scale = float(loss_scale) if loss_scale is not None else (16384.0 if compute_dtype == torch.float16 else 1.0)
        self.fp = FlatParams(ordered_parameters(module), twin_dtype=_twin_dtype_for(module, compute_dtype))
        self.exp_avg = torch.zeros_like(self.fp.flat)
        self.exp_avg_sq = torch.zeros_like(self.fp.flat)
        self.norm_ws = torch.zeros(1, device=self.fp.flat.device, dtype=torch.float32)
        self.d_model, self.n_warmup_steps, self.max_grad_norm = d_model, n_warmup_steps, max_grad_norm
        self.betas, self.eps = betas, eps
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.global_step = 0
        self.lr = 0.0
        self._grads_clean = False
        self.buckets = GradBuckets(self.fp, int(bucket_mb * (1 << 20) / 4))
        self.overlap = bool(overlap) and self.world > 1
        self.lib_collective = LibraryCollective(self.fp.flat.device, process_group) if (collective == "library" and self.world > 1) else None
        self.early_launches = 0                          # buckets whose all-reduce started under backward (diagnostic)
        if self.overlap:
            for s in self.fp.sinks:
                s.owner = self._on_grads_ready        # per-sink notification: several trainers can coexist
                s.no_accumulate = True                # early bucket reductions and gradient accumulation do not mix

    def _reduce(self, lo: int, hi: int, async_op: bool):
        if self.lib_collective is not None:
            return self.lib_collective.all_reduce(self.fp.grad[lo:hi], async_op)
        return dist.all_reduce(self.fp.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def _on_grads_ready(self, sinks) -> None:
        """Called inside backward after an operator enqueued the kernels writing `sinks`: start the all-reduce of every
        bucket that became final.  Every rank runs the same graph in the same order, so the collectives line up."""
        for b in self.buckets.mark_done(sinks):
            b.work = self._reduce(b.lo, b.hi, async_op=True)   # NCCL's stream waits for the kernels enqueued so far
            self.early_launches += 1

    # --- train_multi.py:176-177
    def broadcast_parameters(self, src: int = 0) -> None:
        """Parameters AND optimizer state from rank `src` (hvd.broadcast_parameters + broadcast_optimizer_state,
        train_multi.py:176-177): the Adam moments and the step counter that drives the Noam rate and the bias correction."""
        if self.world > 1:
            dist.broadcast(self.fp.flat, src=src, group=self.group)
            dist.broadcast(self.exp_avg, src=src, group=self.group)
            dist.broadcast(self.exp_avg_sq, src=src, group=self.group)
            step = torch.tensor([self.global_step], dtype=torch.int64, device=self.fp.flat.device)
            dist.broadcast(step, src=src, group=self.group)
            self.global_step = int(step.item())
            self.fp.refresh_rounded()

    # --- Utils.save_model (Utils.py:116-124) writes {'model', 'optimizer'}; train.py:110-114 loads them back
    def state_dict(self) -> dict:
        """Optimizer state in torch.optim.Adam's per-parameter format ('state': {index: {'step', 'exp_avg', 'exp_avg_sq'}},
        'param_groups') — what the reference's ScheduledOptim.optimizer.state_dict() holds — plus the global step that the
        reference forgets to restore.  Parameter indices follow module.parameters() order, as torch.optim does."""
        index = {id(p): i for i, p in enumerate(self.module.parameters())}
        state = {}
        for p, o in zip(self.fp.params, self.fp.offsets):
            n = p.numel()
            state[index[id(p)]] = {"step": torch.tensor(float(self.global_step)),
                                   "exp_avg": self.exp_avg[o:o + n].view_as(p).detach().clone(),
                                   "exp_avg_sq": self.exp_avg_sq[o:o + n].view_as(p).detach().clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "params": sorted(state)}
        return {"state": state, "param_groups": [group], "global_step": self.global_step, "loss_scale": self.loss_scale}

    def load_state_dict(self, sd: dict) -> None:
        index = {id(p): i for i, p in enumerate(self.module.parameters())}
        steps = []
        with torch.no_grad():
            for p, o in zip(self.fp.params, self.fp.offsets):
                st = sd["state"].get(index[id(p)])
                if st is None:
                    continue
                n = p.numel()
                self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps.append(int(float(st.get("step", 0))))
        self.global_step = int(sd.get("global_step", max(steps) if steps else 0))
        if "loss_scale" in sd:
            self.loss_scale = float(sd["loss_scale"])
        self.fp.refresh_rounded()       # the module's parameters were (presumably) just loaded too

    def zero_grad(self) -> None:
        self.fp.zero_grad()
        self.buckets.reset()
        self._grads_clean = False                         # only train_step may vouch for a clean buffer

    # --- train_multi.py:161-163 (the whole model, in flat-buffer order)
    def allreduce_gradients(self):
        """After backward: reduce what the overlap path has not started (everything when overlap is off: ONE collective
        over the whole buffer), then make the compute stream wait for the early collectives."""
        if self.world > 1:
            for lo, hi in self.buckets.pending_ranges():
                self._reduce(lo, hi, async_op=False)
            for b in self.buckets.items:
                if b.work is not None:
                    b.work.wait()
                    b.work = None
        return None

    # --- train.py:45-46
    def step(self) -> None:
        from . import _lib
        from .functional import _stream
        lib = _lib.load()
        self.global_step += 1
        self.lr = noam_lr(self.d_model, self.n_warmup_steps, self.global_step)
        g = self.fp.grad
        self.norm_ws.zero_()
        s = _stream()
        _lib.check(lib.st_sumsq(g.data_ptr(), g.numel(), self.norm_ws.data_ptr(), s))
        a = _lib.AdamArgs(param=self.fp.flat.data_ptr(), grad=g.data_ptr(), exp_avg=self.exp_avg.data_ptr(),
                          exp_avg_sq=self.exp_avg_sq.data_ptr(), n=g.numel(), lr=self.lr, beta1=self.betas[0],
                          beta2=self.betas[1], eps=self.eps, step=self.global_step, max_grad_norm=self.max_grad_norm,
                          grad_scale=1.0 / (self.world * self.loss_scale), norm_ws=self.norm_ws.data_ptr(),
                          param_tf32=None if self.fp.flat_tf32 is None else self.fp.flat_tf32.data_ptr(),
                          twin_dtype=_lib.DTYPE_F32 if self.fp.flat_tf32 is None else
                          {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}[self.fp.flat_tf32.dtype])
        _lib.check(lib.st_adam_step(C.byref(a), s))

    def train_step(self, loss_fn) -> torch.Tensor:
        """One full step: loss_fn() must run forward and return the scalar loss.  The gradient buffer is cleared for the
        NEXT step right after the optimizer kernels are enqueued (so the bookkeeping runs while the GPU is busy even in a
        loop that synchronises on the loss every step): parameter .grad reads zero after this call — use zero_grad /
        backward / allreduce_gradients / step directly to inspect gradients."""
        if not self._grads_clean:
            self.zero_grad()
        self._grads_clean = False
        loss = loss_fn()
        (loss if self.loss_scale == 1.0 else loss * self.loss_scale).backward()
        self.allreduce_gradients()
        self.step()
        self.zero_grad()
        self._grads_clean = True
        return loss
