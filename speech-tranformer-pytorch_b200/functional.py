"""torch-facing wrappers over the C ABI (include/st_b200.h): raw-pointer plumbing and autograd.

PyTorch is used for device memory, streams and autograd bookkeeping only; every kernel that runs is
from libst_b200.so.  Nothing here falls back to a PyTorch implementation: CPU tensors, non-fp32
tensors or a missing library raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import StError, check

__all__ = ["add_layer_norm", "multi_head_attention", "positionwise_ffn", "label_smoothing_ce", "soft_target_ce",
           "attention_core", "linear_tf32", "round_tf32", "is_tf32_clean", "mark_tf32_clean", "next_seed",
           "frontend", "linear", "embedding", "ctc_loss", "GradSink", "attach_grad_sink", "attach_tf32_twin",
           "LengthMask", "cast", "ACT_DTYPES", "set_fp32_engine", "fp32_engine"]

# activation element types and their ST_DTYPE_* codes (include/st_b200.h).  fp32 tensors take the fp32 / TF32 path; fp16 and
# bf16 tensors take the 16-bit path (tcgen05 kind::f16 operands, fp32 accumulation and statistics; d_k must be 64).
ACT_DTYPES = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}

# How the composite operators (multi_head_attention, positionwise_ffn) compute on fp32 tensors:
#   "fp16"  ST_DTYPE_F32_H16 — fp32 tensors at the operator boundary (inputs, outputs, every gradient), fp16 operands with
#           fp32 accumulation inside; fp16 carries the same 11 significant bits as TF32, and the backward pass scales its
#           16-bit tensors by a power of two derived on the device from max|grad_output| so they stay in fp16's range.
#           Needs d_k = 64 for attention (other head sizes use "tf32").
#   "tf32"  ST_DTYPE_F32 — TF32 operands everywhere (round 1's path).
# Set with set_fp32_engine() or the environment variable ST_FP32_ENGINE before import.
_FP32_ENGINES = ("fp16", "tf32")
_fp32_engine = [os.environ.get("ST_FP32_ENGINE", "fp16")]
if _fp32_engine[0] not in _FP32_ENGINES:
    raise RuntimeError(f"ST_FP32_ENGINE must be one of {_FP32_ENGINES}, got {_fp32_engine[0]!r}")


def set_fp32_engine(name: str) -> str:
    """Choose how the composite operators compute on fp32 tensors ("fp16" or "tf32"); returns the previous setting."""
    if name not in _FP32_ENGINES:
        raise ValueError(f"fp32 engine must be one of {_FP32_ENGINES}, got {name!r}")
    prev, _fp32_engine[0] = _fp32_engine[0], name
    return prev


def fp32_engine() -> str:
    return _fp32_engine[0]


# ------------------------------------------------------------------------------------------------
# plumbing
# ------------------------------------------------------------------------------------------------
def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor — the B200 hot path has no CPU fallback")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t


def _need_act(t: torch.Tensor, name: str, like: Optional[torch.Tensor] = None) -> torch.Tensor:
    """An activation tensor: fp32, fp16 or bf16 on the device (the same type as `like` when given)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor — the B200 hot path has no CPU fallback")
    if t.dtype not in ACT_DTYPES:
        raise RuntimeError(f"{name}: expected dtype float32, float16 or bfloat16, got {t.dtype}")
    if like is not None and t.dtype != like.dtype:
        raise RuntimeError(f"{name}: expected dtype {like.dtype} like the other activations, got {t.dtype}")
    return t


class LengthMask:
    """A key-padding (+ optional causal) mask given by its lengths instead of a (B, Lq, Lk) tensor: key j of batch b is
    masked iff j >= k_len[b] (Utils.padding_info_mask, Utils.py:41-57) or, when `causal`, j > i (Utils.feature_info_mask,
    Utils.py:60-70, ORed as in Models.py:89-94).  Pass it wherever the modules take `mask`: the kernels derive every
    predicate from the lengths, no mask tensor is built or scanned.  dense() materialises the equivalent bool tensor."""
    __slots__ = ("k_len", "len_q", "len_k", "causal")

    def __init__(self, k_len: torch.Tensor, len_q: int, len_k: int, causal: bool = False):
        if k_len.dtype != torch.int64 or k_len.dim() != 1:
            raise RuntimeError("LengthMask: k_len must be a 1-D int64 tensor")
        self.k_len, self.len_q, self.len_k, self.causal = k_len, int(len_q), int(len_k), bool(causal)

    def size(self):
        return torch.Size((self.k_len.shape[0], self.len_q, self.len_k))

    def dense(self) -> torch.Tensor:
        ar = torch.arange(self.len_k, device=self.k_len.device)
        m = (ar.unsqueeze(0) >= self.k_len.unsqueeze(1)).unsqueeze(1).expand(-1, self.len_q, -1)
        if self.causal:
            m = m | torch.ones(self.len_q, self.len_k, dtype=torch.bool, device=m.device).triu(1).unsqueeze(0)
        return m


def _contig(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


_device_ok = set()


def _lib_for(t: torch.Tensor):
    lib = _lib.load()
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if dev not in _device_ok:
        check(lib.st_device_check(dev))
        _device_ok.add(dev)
    return lib


# Module outputs feed the next module's GEMMs.  True: the producing kernel rounds them to TF32 in its epilogue (no
# extra pass; the residual stream then carries one 2^-11 rounding per module).  False: outputs stay exact fp32 and
# every consumer makes its own rounded operand copy (one extra read+write of the activation per module).
ROUND_OUT = os.environ.get("ST_ROUND_OUT", "1") != "0"

_TF32_TAG = "_st_tf32_clean"


def is_tf32_clean(t: torch.Tensor) -> bool:
    """True if `t` was produced by this library already rounded to TF32 (so no rounding copy is needed)."""
    return bool(getattr(t, _TF32_TAG, False))


def mark_tf32_clean(t: torch.Tensor) -> torch.Tensor:
    setattr(t, _TF32_TAG, True)
    return t


class GradSink:
    """Where a parameter's gradient lives in a trainer-owned flat buffer (parallel.FlatParams).

    The composite backward kernels OVERWRITE their parameter gradients, so when every parameter of an operator
    carries a sink they write straight into the flat gradient buffer and return None to autograd — no temporary, no
    AccumulateGrad `+=` kernel (263 of them per step in the 6+6 model).  `written` guards weight sharing: a second
    use of the same parameter within one step falls back to the ordinary accumulate path.  `zeroed` is set by the owner of
    the buffer (FlatParams.zero_grad) after it cleared the whole buffer in one pass: the backward operators then skip their
    own per-tensor clears (grads_zeroed in the st_*_bwd_args; ~200 memset nodes per training step otherwise)."""
    __slots__ = ("view", "written", "zeroed", "uses", "done", "owner", "no_accumulate")

    def __init__(self, view: torch.Tensor):
        self.view, self.written, self.zeroed = view, False, False
        self.owner = None      # callable(sinks) of the buffer's owner, told when backward operators have written slices
        # set by an owner that all-reduces slices DURING backward: a second forward before zero_grad would then add local
        # gradients on top of already-reduced sums (ranks diverge silently), so it is refused
        self.no_accumulate = False
        # uses: forward operators that took this parameter since the last zero_grad; done: backward operators that have
        # since written its gradient directly (kernels enqueued).  done == uses >= 1 means the slice is final for this step
        # (stream order) — what parallel.DataParallelTrainer needs to start a bucket's all-reduce under the backward pass.
        self.uses, self.done = 0, 0


_SINK_ATTR = "_st_grad_sink"
_TWIN_ATTR = "_st_tf32_twin"


_stale_twin_warned = [False]


def attach_tf32_twin(param: torch.Tensor, view: torch.Tensor) -> None:
    """Register a caller-maintained operand-precision copy of `param` — TF32-rounded fp32, or fp16 / bf16 for the 16-bit
    path (parallel.FlatParams keeps it current from inside the Adam kernel).  The copy is used only while the
    parameter's version counter is unchanged, so any PyTorch-side in-place modification (load_state_dict, init, ...)
    falls back to rounding the weight per call (logged once: it is a performance cliff, not an error)."""
    setattr(param, _TWIN_ATTR, (view, param._version))


def _twins_of(params, dtype=torch.float32):
    """The params' operand-precision twins of element type `dtype` if ALL of them have a current one, else None."""
    out = []
    for p in params:
        tw = getattr(p, _TWIN_ATTR, None)
        if tw is None or tw[0].dtype != dtype or tw[0].shape != p.shape:
            return None
        if tw[1] != p._version:
            if not _stale_twin_warned[0]:
                _stale_twin_warned[0] = True
                import warnings
                warnings.warn("speech-tranformer-pytorch_b200: a parameter with a registered operand-precision twin was modified "
                              "in place; its weights are rounded per call until FlatParams.refresh_rounded() is called")
            return None
        out.append(tw[0])
    return out


def _composite_dtype(act_dtype, dk=None):
    """(ST_DTYPE_* code, element type of the weight twins) a composite operator runs in for activations of `act_dtype`."""
    if act_dtype != torch.float32:
        return ACT_DTYPES[act_dtype], act_dtype
    if _fp32_engine[0] == "fp16" and (dk is None or dk == 64):
        return _lib.DTYPE_F32_H16, torch.float16
    return _lib.DTYPE_F32, torch.float32


# Chaining of the fp16-operand engine (ST_DTYPE_F32_H16).  An fp32 activation produced by one composite operator carries the
# fp16 copy its LayerNorm kernel wrote next to it, and an fp32 gradient produced by one backward operator carries the device
# scalar max|gradient| its last GEMM epilogue measured; the next operator finds them here and skips its conversion pass /
# its amax pass.  A tag is honoured only for the very tensor object it was attached to and only while that tensor's version
# counter is unchanged (an in-place edit, or autograd accumulating another gradient into it, invalidates it).  Writes
# through `tensor.data` bypass the version counter — as everywhere in PyTorch, they are invisible to such caches; use
# ST_CHAIN=0 if a training loop edits activations that way.
_H16_ATTR = "_st_h16"
_AMAX_ATTR = "_st_amax"
_CHAIN = os.environ.get("ST_CHAIN", "1") != "0"      # ST_CHAIN=0: attach no tags (A/B measurements)


def _tag(t: torch.Tensor, attr: str, value: torch.Tensor) -> None:
    if _CHAIN:
        setattr(t, attr, (value, t._version))


def _tagged(t: torch.Tensor, attr: str) -> Optional[torch.Tensor]:
    tag = getattr(t, attr, None)
    if tag is None or tag[1] != t._version:
        return None
    return tag[0]


def _h16_of(t: torch.Tensor) -> torch.Tensor:
    """The fp16 copy of the contiguous fp32 activation `t`: the one its producer left, else made (and remembered) here."""
    tw = _tagged(t, _H16_ATTR)
    if tw is None or tw.shape != t.shape:
        tw = cast(t, torch.float16)
        _tag(t, _H16_ATTR, tw)
    return tw


def attach_grad_sink(param: torch.Tensor, view: torch.Tensor) -> GradSink:
    sink = GradSink(view)
    setattr(param, _SINK_ATTR, sink)
    return sink


def _sinks_of(params, ctx=None):
    """The params' sinks if ALL of them have an unwritten sink of the right shape, else None.  Called from an operator's
    forward with its autograd context; a forward that will never see a backward (torch.no_grad(), eval / decode passes:
    ctx.needs_input_grad is all False) takes no sinks and counts no use, so the use / done bookkeeping that
    parallel.GradBuckets relies on stays balanced across ranks."""
    if ctx is not None and not any(ctx.needs_input_grad):
        return None
    found = [None if p is None else getattr(p, _SINK_ATTR, None) for p in params]
    for s in found:
        if s is not None:
            if s.written and s.no_accumulate:
                raise RuntimeError("gradient accumulation (a second forward/backward before zero_grad) is not supported while "
                                   "the trainer all-reduces gradient buckets under the backward pass: construct "
                                   "DataParallelTrainer(overlap=False) to accumulate")
            s.uses += 1
    out = []
    for p, s in zip(params, found):
        if p is None:
            out.append(None)
            continue
        if s is None or s.written or s.view.shape != p.shape or not s.view.is_contiguous():
            return None
        out.append(s)
    return out


def _notify(sinks, direct) -> None:
    """Inside backward, right after an operator has ENQUEUED the kernels that write these sinks' gradients: count the
    write and tell each sink's owner (parallel.DataParallelTrainer uses this to start a bucket's all-reduce under the
    rest of the backward pass).  Nothing happens on the autograd-accumulate fallback path (direct is None)."""
    if direct is None or not sinks:
        return
    owners = {}
    for s in sinks:
        if s is not None:
            s.done += 1
            if s.owner is not None:
                owners.setdefault(id(s.owner), (s.owner, []))[1].append(s)
    for cb, group in owners.values():
        cb(group)


def _claim(sinks):
    """Inside backward: take the sinks (None if any was written meanwhile) and mark them written.  Returns
    (views or None, zeroed): zeroed = 1 when every view is known to hold zeros (see GradSink)."""
    if sinks is None or any(s is not None and s.written for s in sinks):
        return None, 0
    zeroed = 1
    for s in sinks:
        if s is not None:
            zeroed &= int(s.zeroed)
            s.written, s.zeroed = True, False
    return [None if s is None else s.view for s in sinks], zeroed


_seed_state = [0]


def next_seed() -> int:
    """A fresh 63-bit dropout seed derived from torch's CPU generator (reproducible under manual_seed)."""
    _seed_state[0] += 1
    base = int(torch.empty((), dtype=torch.int64).random_().item())
    return (base ^ (_seed_state[0] * 0x9E3779B97F4A7C15)) & 0x7FFFFFFFFFFFFFFF


def _same(a: torch.Tensor, b: torch.Tensor) -> bool:
    return a is b or (a.data_ptr() == b.data_ptr() and a.shape == b.shape and a.stride() == b.stride())


def _split_mask(mask, B: int, Lq: int, Lk: int, device):
    """(tensor mask or None, k_len or None, causal) from what a module received as `mask`."""
    if isinstance(mask, LengthMask):
        if tuple(mask.size()) != (B, Lq, Lk):
            raise RuntimeError(f"mask: expected shape {(B, Lq, Lk)}, got {tuple(mask.size())}")
        k_len = mask.k_len if mask.k_len.device == device else mask.k_len.to(device)
        return None, _contig(k_len), int(mask.causal)
    return mask, None, 0


def _mask_args(mask: Optional[torch.Tensor], B: int, Lq: int, Lk: int, device) -> Tuple[Optional[torch.Tensor], int, int, int]:
    """Accept bool or uint8, any strides (stride-0 broadcast views are NOT materialised)."""
    if mask is None:
        return None, 0, 0, 0
    if mask.dtype == torch.bool:
        mask = mask.view(torch.uint8)
    elif mask.dtype != torch.uint8:
        raise RuntimeError(f"mask: expected bool or uint8, got {mask.dtype}")
    if tuple(mask.shape) != (B, Lq, Lk):
        raise RuntimeError(f"mask: expected shape {(B, Lq, Lk)}, got {tuple(mask.shape)}")
    if mask.device != device:
        mask = mask.to(device)
    sb, sq, sk = mask.stride()
    return mask, sb, sq, sk


# ------------------------------------------------------------------------------------------------
# elementwise helpers
# ------------------------------------------------------------------------------------------------
def round_tf32(x: torch.Tensor) -> torch.Tensor:
    x = _contig(_need(x, "x"))
    lib = _lib_for(x)
    out = torch.empty_like(x)
    cols = x.shape[-1]
    rows = x.numel() // max(cols, 1)
    check(lib.st_round_tf32(_p(x), cols, _p(out), cols, rows, cols, _stream()))
    return mark_tf32_clean(out)


def cast(x: torch.Tensor, dtype, scale: float = 1.0) -> torch.Tensor:
    """x converted to `dtype` (fp32 <-> fp16 / bf16) by the library's conversion kernel, optionally scaled (no autograd)."""
    x = _contig(_need_act(x, "x"))
    if dtype not in ACT_DTYPES:
        raise RuntimeError(f"cast: dtype must be float32, float16 or bfloat16, got {dtype}")
    lib = _lib_for(x)
    out = torch.empty_like(x, dtype=dtype)
    cols = x.shape[-1] if x.dim() else 1
    rows = x.numel() // max(cols, 1)
    check(lib.st_cast(_p(x), ACT_DTYPES[x.dtype], cols, _p(out), ACT_DTYPES[dtype], cols, rows, cols, float(scale), _stream()))
    return out


def linear_tf32(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False
                ) -> torch.Tensor:
    """y = x @ weight.T + bias on the tcgen05 TF32 GEMM (no autograd). Operands are rounded to TF32 first."""
    x = _contig(_need(x, "x"))
    weight = _contig(_need(weight, "weight"))
    lib = _lib_for(x)
    K = x.shape[-1]
    M = x.numel() // K
    N = weight.shape[0]
    xr = x if is_tf32_clean(x) else round_tf32(x)
    wr = round_tf32(weight)
    out = torch.empty(*x.shape[:-1], N, device=x.device, dtype=torch.float32)
    ep = _lib.GemmEpilogue(bias=_p(bias), aux=None, ldaux=0, aux_mode=0, relu=int(relu), round_tf32=0, k_splits=1,
                           dropout_p=0.0, seed=0)
    check(lib.st_gemm(0, _p(xr), K, _p(wr), K, _p(out), N, M, N, K, C.byref(ep), _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# (C) residual + LayerNorm
# ------------------------------------------------------------------------------------------------
class _AddLayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, gamma, beta, eps, dropout_p, seed, round_out):
        a = _contig(_need(a, "a"))
        lib = _lib_for(a)
        d = a.shape[-1]
        rows = a.numel() // d
        if b is not None:
            b = _contig(_need(b, "b"))
            if b.shape != a.shape:
                raise RuntimeError(f"add_layer_norm: shapes differ {tuple(a.shape)} vs {tuple(b.shape)}")
        gamma = _contig(_need(gamma, "gamma"))
        beta = _contig(_need(beta, "beta"))
        out = torch.empty_like(a)
        z = torch.empty_like(a) if b is not None else None
        mean = torch.empty(rows, device=a.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=a.device, dtype=torch.float32)
        check(lib.st_add_ln_fwd(_p(a), _p(b), _p(gamma), _p(beta), _p(out), _p(z), _p(mean), _p(rstd), rows, d,
                                float(eps), int(round_out), float(dropout_p), int(seed), _stream()))
        ctx.save_for_backward(z if z is not None else a, mean, rstd, gamma)
        ctx.cfg = (rows, d, float(dropout_p), int(seed), b is not None)
        if round_out:
            mark_tf32_clean(out)
        return out

    @staticmethod
    def backward(ctx, dy):
        z, mean, rstd, gamma = ctx.saved_tensors
        rows, d, p, seed, has_b = ctx.cfg
        dy = _contig(dy)
        lib = _lib_for(dy)
        dz = torch.empty_like(z)
        dgamma = torch.zeros(d, device=z.device, dtype=torch.float32)
        dbeta = torch.zeros(d, device=z.device, dtype=torch.float32)
        check(lib.st_add_ln_bwd(_p(dy), _p(z), _p(mean), _p(rstd), _p(gamma), _p(dz), _p(dgamma), _p(dbeta), None,
                                rows, d, 0, p, seed, _stream()))
        return dz, (dz if has_b else None), dgamma, dbeta, None, None, None, None


def add_layer_norm(a, b, gamma, beta, eps: float = 1e-6, dropout_p: float = 0.0, seed: int = 0,
                   round_out: bool = False):
    """dropout(LayerNorm(a + b) * gamma + beta) — Attention.py:94, SubLayers.py:27."""
    return _AddLayerNorm.apply(a, b, gamma, beta, eps, dropout_p, seed, round_out)


# ------------------------------------------------------------------------------------------------
# (A) attention core (no projections) — also ScaledDotProductAttention
# ------------------------------------------------------------------------------------------------
class _AttentionCore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, mask, n_head, dropout_p, seed, need_attn):
        ctx.set_materialize_grads(False)   # the auxiliary outputs never carry a gradient: no zero tensors for them
        q = _contig(_need_act(q, "q"))
        k, v = _contig(_need_act(k, "k", q)), _contig(_need_act(v, "v", q))
        lib = _lib_for(q)
        dt = ACT_DTYPES[q.dtype]
        B, Lq, d = q.shape
        Lk = k.shape[1]
        dk = d // n_head
        mask, k_len, causal = _split_mask(mask, B, Lq, Lk, q.device)
        mask_t, sb, sq, sk = _mask_args(mask, B, Lq, Lk, q.device)
        if dt == _lib.DTYPE_F32:
            qr = q if is_tf32_clean(q) else round_tf32(q)
            kr = qr if _same(k, q) else (k if is_tf32_clean(k) else round_tf32(k))
            vr = kr if _same(v, k) else (v if is_tf32_clean(v) else round_tf32(v))
        else:
            qr, kr, vr = q, k, v
        out = torch.empty(B, Lq, d, device=q.device, dtype=q.dtype)
        lse = torch.empty(B, n_head, Lq, device=q.device, dtype=torch.float32)
        attn = torch.empty(B, n_head, Lq, Lk, device=q.device, dtype=torch.float32) if need_attn else None
        a = _lib.AttnArgs(B=B, H=n_head, Lq=Lq, Lk=Lk, dk=dk, q=_p(qr), ldq=d, k=_p(kr), ldk=d, v=_p(vr), ldv=d,
                          mask=_p(mask_t), ms_b=sb, ms_q=sq, ms_k=sk, dropout_p=float(dropout_p), seed=int(seed),
                          ctx=_p(out), ldctx=d, lse=_p(lse), attn=_p(attn), dtype=dt, k_len=_p(k_len), causal=causal)
        check(lib.st_attn_fwd(C.byref(a), _stream()))
        ctx.save_for_backward(qr, kr, vr, out, lse, mask_t, k_len)
        ctx.cfg = (B, n_head, Lq, Lk, dk, float(dropout_p), int(seed), dt, causal)
        ctx.mark_non_differentiable(*([attn] if attn is not None else []))
        return out, attn

    @staticmethod
    def backward(ctx, dout, _dattn):
        if dout is None:      # only the auxiliary output was used downstream: no gradient flows
            return (None,) * 8
        qr, kr, vr, out, lse, mask_t, k_len = ctx.saved_tensors
        B, H, Lq, Lk, dk, p, seed, dt, causal = ctx.cfg
        d = H * dk
        lib = _lib_for(qr)
        dor = round_tf32(_contig(dout)) if dt == _lib.DTYPE_F32 else _contig(_need_act(dout, "grad_output", qr))
        _, sb, sq, sk = _mask_args(mask_t, B, Lq, Lk, qr.device)
        dq = torch.empty_like(qr)
        dkk = torch.empty_like(kr)
        dv = torch.empty_like(vr)
        delta = torch.empty(B, H, Lq, device=qr.device, dtype=torch.float32)
        f = _lib.AttnArgs(B=B, H=H, Lq=Lq, Lk=Lk, dk=dk, q=_p(qr), ldq=d, k=_p(kr), ldk=d, v=_p(vr), ldv=d,
                          mask=_p(mask_t), ms_b=sb, ms_q=sq, ms_k=sk, dropout_p=p, seed=seed,
                          ctx=_p(out), ldctx=d, lse=_p(lse), attn=None, dtype=dt, k_len=_p(k_len), causal=causal)
        a = _lib.AttnBwdArgs(f=f, dctx=_p(dor), lddctx=d, delta=_p(delta), dq=_p(dq), lddq=d, dk=_p(dkk), lddk=d,
                             dv=_p(dv), lddv=d)
        check(lib.st_attn_bwd(C.byref(a), _stream()))
        return dq, dkk, dv, None, None, None, None, None


def attention_core(q, k, v, mask=None, n_head: int = 1, dropout_p: float = 0.0, seed: int = 0,
                   need_attn: bool = False):
    """softmax(mask(q k^T / sqrt(d_k))) v over `n_head` heads; q,k,v are (B, L, n_head*d_k)."""
    return _AttentionCore.apply(q, k, v, mask, n_head, dropout_p, seed, need_attn)


# ------------------------------------------------------------------------------------------------
# composite MultiHeadAttention
# ------------------------------------------------------------------------------------------------
class _MultiHeadAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, mask, wq, bq, wk, bk, wv, bv, wo, bo, ln_g, ln_b, n_head, residual, eps, dropout_p,
                seed, need_attn, round_out):
        ctx.set_materialize_grads(False)   # the auxiliary outputs never carry a gradient: no zero tensors for them
        q = _need_act(q, "q")
        k, v = _need_act(k, "k", q), _need_act(v, "v", q)
        same_qk, same_kv = _same(q, k), _same(k, v)
        qc = _contig(q)
        kc = qc if same_qk else _contig(k)
        vc = kc if same_kv else _contig(v)
        lib = _lib_for(qc)
        if qc.dim() != 3 or kc.dim() != 3 or vc.dim() != 3:
            raise RuntimeError("multi_head_attention: q, k, v must be (batch, length, d_model)")
        B, Lq, d = qc.shape
        Lk = kc.shape[1]
        if vc.shape != kc.shape or kc.shape[0] != B or kc.shape[2] != d:
            raise RuntimeError(f"multi_head_attention: incompatible shapes q{tuple(qc.shape)} k{tuple(kc.shape)} v{tuple(vc.shape)}")
        dk = d // n_head
        dt, twin_dtype = _composite_dtype(qc.dtype, dk)
        res = vc if residual == "v" else qc
        if res.shape != qc.shape:
            # the reference's `output + v` (Attention.py:94) fails the same way when len_q != len_k
            raise RuntimeError(f"The size of tensor a ({Lq}) must match the size of tensor b ({res.shape[1]}) at "
                               "non-singleton dimension 1 (residual='v' with len_q != len_k; use residual='q')")
        mask, k_len, causal = _split_mask(mask, B, Lq, Lk, qc.device)
        mask_t, sb, sq, sk = _mask_args(mask, B, Lq, Lk, qc.device)
        params = [_contig(_need(t, "parameter")) for t in (wq, bq, wk, bk, wv, bv, wo, bo, ln_g, ln_b)]
        inputs_tf32 = int(dt != _lib.DTYPE_F32 or (is_tf32_clean(q) and is_tf32_clean(k) and is_tf32_clean(v)))
        q16 = k16 = v16 = out16 = None
        if dt == _lib.DTYPE_F32_H16:      # fp16 operand copies: the producers' (or made once per tensor), see _h16_of
            q16 = _h16_of(qc)
            k16 = q16 if same_qk else _h16_of(kc)
            v16 = k16 if same_kv else _h16_of(vc)
            out16 = torch.empty(B, Lq, d, device=qc.device, dtype=torch.float16)
        same_qkv = int(same_qk and same_kv and Lq == Lk)
        n_saved = lib.st_mha_saved_floats_dt(dt, B, Lq, Lk, n_head, d, same_qkv, int(same_kv), inputs_tf32)
        saved = torch.empty(n_saved, device=qc.device, dtype=torch.float32)
        out = torch.empty(B, Lq, d, device=qc.device, dtype=qc.dtype)
        attn = torch.empty(B, n_head, Lq, Lk, device=qc.device, dtype=torch.float32) if need_attn else None
        twins = _twins_of((wq, wk, wv, wo), twin_dtype) or [None] * 4
        a = _lib.MhaArgs(B=B, Lq=Lq, Lk=Lk, H=n_head, d_model=d, dk=dk, q_in=_p(qc), k_in=_p(kc), v_in=_p(vc),
                         residual=_p(res), wq=_p(params[0]), bq=_p(params[1]), wk=_p(params[2]), bk=_p(params[3]),
                         wv=_p(params[4]), bv=_p(params[5]), wo=_p(params[6]), bo=_p(params[7]), ln_g=_p(params[8]),
                         ln_b=_p(params[9]), mask=_p(mask_t), ms_b=sb, ms_q=sq, ms_k=sk, eps=float(eps),
                         dropout_p=float(dropout_p), seed=int(seed), inputs_tf32=inputs_tf32, round_out=int(round_out),
                         out=_p(out), attn=_p(attn), saved=_p(saved), saved_floats=n_saved, ws=None, ws_floats=0,
                         wq_tf32=_p(twins[0]), wk_tf32=_p(twins[1]), wv_tf32=_p(twins[2]), wo_tf32=_p(twins[3]),
                         dtype=dt, k_len=_p(k_len), causal=causal, q_h16=_p(q16), k_h16=_p(k16), v_h16=_p(v16),
                         out_h16=_p(out16))
        check(lib.st_mha_fwd(C.byref(a), _stream()))
        ctx.twins = twins          # backward must present the same weight copies the forward used
        ctx.h16 = (q16, k16, v16)  # ... and the same operand copies of the inputs
        if out16 is not None:
            _tag(out, _H16_ATTR, out16)
        ctx.save_for_backward(qc, kc, vc, mask_t, k_len, saved, *params)
        ctx.sinks = _sinks_of((wq, bq, wk, bk, wv, bv, wo, bo, ln_g, ln_b), ctx)
        ctx.cfg = (B, Lq, Lk, n_head, d, dk, residual, float(eps), float(dropout_p), int(seed), inputs_tf32,
                   same_qk, same_kv, dt, causal)
        if attn is not None:
            ctx.mark_non_differentiable(attn)
        if round_out and dt == _lib.DTYPE_F32:
            mark_tf32_clean(out)
        return out, attn

    @staticmethod
    def backward(ctx, dout, _dattn):
        if dout is None:      # only the auxiliary output was used downstream: no gradient flows
            return (None,) * 21
        qc, kc, vc, mask_t, k_len, saved, *params = ctx.saved_tensors
        B, Lq, Lk, H, d, dk, residual, eps, p, seed, inputs_tf32, same_qk, same_kv, dt, causal = ctx.cfg
        lib = _lib_for(qc)
        dout = _contig(_need_act(dout, "grad_output", qc))
        dev = qc.device
        _, sb, sq, sk = _mask_args(mask_t, B, Lq, Lk, dev)
        res = vc if residual == "v" else qc
        n_ws = lib.st_mha_ws_floats_dt(dt, B, Lq, Lk, H, d)
        ws = torch.empty(n_ws, device=dev, dtype=torch.float32)
        same_qkv = same_qk and same_kv
        dq_in = torch.empty_like(qc)
        dk_in = dq_in if same_qkv else torch.empty_like(kc)
        dv_in = dk_in if same_kv else torch.empty_like(vc)
        direct, zeroed = _claim(ctx.sinks)
        grads = direct if direct is not None else [torch.empty_like(t) for t in params]
        if direct is None and params[0].shape == params[2].shape == params[4].shape == (d, d):
            # dW / db of the three projections as slices of one packed buffer: st_mha_bwd then computes projections that
            # share an input (self-attention: all three; cross-attention: k and v) with ONE wgrad GEMM / column sum
            gw = torch.empty(3 * d, d, device=dev, dtype=torch.float32)
            gb = torch.empty(3 * d, device=dev, dtype=torch.float32)
            for i in range(3):
                grads[2 * i], grads[2 * i + 1] = gw[i * d:(i + 1) * d], gb[i * d:(i + 1) * d]
        f = _lib.MhaArgs(B=B, Lq=Lq, Lk=Lk, H=H, d_model=d, dk=dk, q_in=_p(qc), k_in=_p(kc), v_in=_p(vc),
                         residual=_p(res), wq=_p(params[0]), bq=_p(params[1]), wk=_p(params[2]), bk=_p(params[3]),
                         wv=_p(params[4]), bv=_p(params[5]), wo=_p(params[6]), bo=_p(params[7]), ln_g=_p(params[8]),
                         ln_b=_p(params[9]), mask=_p(mask_t), ms_b=sb, ms_q=sq, ms_k=sk, eps=eps, dropout_p=p,
                         seed=seed, inputs_tf32=inputs_tf32, round_out=0, out=None, attn=None, saved=_p(saved),
                         saved_floats=saved.numel(), ws=_p(ws), ws_floats=n_ws, wq_tf32=_p(ctx.twins[0]),
                         wk_tf32=_p(ctx.twins[1]), wv_tf32=_p(ctx.twins[2]), wo_tf32=_p(ctx.twins[3]),
                         dtype=dt, k_len=_p(k_len), causal=causal, q_h16=_p(ctx.h16[0]), k_h16=_p(ctx.h16[1]),
                         v_h16=_p(ctx.h16[2]), out_h16=None)
        mixed = dt == _lib.DTYPE_F32_H16
        dq_amax = torch.empty(1, device=dev, dtype=torch.float32) if mixed else None    # cleared by the operator
        a = _lib.MhaBwdArgs(f=f, dout=_p(dout), dq_in=_p(dq_in), dk_in=_p(dk_in), dv_in=_p(dv_in), dresidual=None,
                            dwq=_p(grads[0]), dbq=_p(grads[1]), dwk=_p(grads[2]), dbk=_p(grads[3]), dwv=_p(grads[4]),
                            dbv=_p(grads[5]), dwo=_p(grads[6]), dbo=_p(grads[7]), dln_g=_p(grads[8]),
                            dln_b=_p(grads[9]), grads_zeroed=zeroed,
                            dout_amax=_p(_tagged(dout, _AMAX_ATTR)) if mixed else None, dq_amax=_p(dq_amax))
        check(lib.st_mha_bwd(C.byref(a), _stream()))
        if mixed:
            _tag(dq_in, _AMAX_ATTR, dq_amax)
        _notify(ctx.sinks, direct)
        # aliased inputs received ONE combined gradient; hand it to the first alias only
        gq, gk, gv = dq_in, (None if same_qkv else dk_in), (None if same_kv else dv_in)
        if same_qk and not same_kv:   # q is k but v differs: separate buffers were filled, combine them
            gq, gk = dq_in + dk_in, None
        if direct is not None:        # parameter gradients already sit in the trainer's flat buffer
            grads = [None] * len(params)
        return (gq, gk, gv, None, *grads, None, None, None, None, None, None, None)


def multi_head_attention(q, k, v, mask, wq, bq, wk, bk, wv, bv, wo, bo, ln_g, ln_b, n_head: int, residual: str = "v",
                         eps: float = 1e-6, dropout_p: float = 0.0, seed: int = 0, need_attn: bool = False,
                         round_out: bool = True):
    """MultiHeadAttention.forward (Attention.py:64-96) as one fused operator. Returns (out, attn or None)."""
    if residual not in ("v", "q"):
        raise ValueError("residual must be 'v' (reference behaviour) or 'q'")
    return _MultiHeadAttention.apply(q, k, v, mask, wq, bq, wk, bk, wv, bv, wo, bo, ln_g, ln_b, n_head, residual, eps,
                                     dropout_p, seed, need_attn, round_out)


# ------------------------------------------------------------------------------------------------
# composite PositionwiseFeedForward
# ------------------------------------------------------------------------------------------------
class _PositionwiseFFN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, ln_g, ln_b, eps, dropout_p, seed, round_out):
        ctx.set_materialize_grads(False)   # the auxiliary outputs never carry a gradient: no zero tensors for them
        xc = _contig(_need_act(x, "inputs"))
        dt, twin_dtype = _composite_dtype(xc.dtype)
        x_clean = int(dt != _lib.DTYPE_F32 or is_tf32_clean(x))
        lib = _lib_for(xc)
        d = xc.shape[-1]
        rows = xc.numel() // d
        params = [_contig(_need(t, "parameter")) for t in (w1, b1, w2, b2, ln_g, ln_b)]
        d_ff = params[0].shape[0]
        x16 = out16 = None
        if dt == _lib.DTYPE_F32_H16:
            x16 = _h16_of(xc)
            out16 = torch.empty_like(xc, dtype=torch.float16)
        n_saved = lib.st_ffn_saved_floats_dt(dt, rows, d, d_ff, x_clean)
        saved = torch.empty(n_saved, device=xc.device, dtype=torch.float32)
        out = torch.empty_like(xc)
        twins = _twins_of((w1, w2), twin_dtype) or [None] * 2
        a = _lib.FfnArgs(rows=rows, d_model=d, d_ff=d_ff, x=_p(xc), w1=_p(params[0]), b1=_p(params[1]),
                         w2=_p(params[2]), b2=_p(params[3]), ln_g=_p(params[4]), ln_b=_p(params[5]), eps=float(eps),
                         dropout_p=float(dropout_p), seed=int(seed), x_is_tf32=x_clean, round_out=int(round_out),
                         out=_p(out), saved=_p(saved), saved_floats=n_saved, ws=None, ws_floats=0,
                         w1_tf32=_p(twins[0]), w2_tf32=_p(twins[1]), dtype=dt, x_h16=_p(x16), out_h16=_p(out16))
        check(lib.st_ffn_fwd(C.byref(a), _stream()))
        ctx.twins = twins
        ctx.x16 = x16
        if out16 is not None:
            _tag(out, _H16_ATTR, out16)
        ctx.save_for_backward(xc, saved, *params)
        ctx.sinks = _sinks_of((w1, b1, w2, b2, ln_g, ln_b), ctx)
        ctx.cfg = (rows, d, d_ff, float(eps), float(dropout_p), int(seed), x_clean, dt)
        if round_out and dt == _lib.DTYPE_F32:
            mark_tf32_clean(out)
        off = lib.st_ffn_hidden_offset_dt(dt, rows, d, d_ff, x_clean)   # the hidden activation follows the optional input copy
        if dt == _lib.DTYPE_F32:
            hidden = saved[off:off + rows * d_ff].view(*xc.shape[:-1], d_ff)
        else:
            hidden = saved.view(twin_dtype)[2 * off:2 * off + rows * d_ff].view(*xc.shape[:-1], d_ff)
        ctx.mark_non_differentiable(hidden)
        return out, hidden

    @staticmethod
    def backward(ctx, dout, _dhidden):
        if dout is None:      # only the auxiliary output was used downstream: no gradient flows
            return (None,) * 11
        xc, saved, *params = ctx.saved_tensors
        rows, d, d_ff, eps, p, seed, x_clean, dt = ctx.cfg
        lib = _lib_for(xc)
        dout = _contig(_need_act(dout, "grad_output", xc))
        n_ws = lib.st_ffn_ws_floats_dt(dt, rows, d, d_ff)
        ws = torch.empty(n_ws, device=xc.device, dtype=torch.float32)
        dx = torch.empty_like(xc)
        direct, zeroed = _claim(ctx.sinks)
        grads = direct if direct is not None else [torch.empty_like(t) for t in params]
        f = _lib.FfnArgs(rows=rows, d_model=d, d_ff=d_ff, x=_p(xc), w1=_p(params[0]), b1=_p(params[1]),
                         w2=_p(params[2]), b2=_p(params[3]), ln_g=_p(params[4]), ln_b=_p(params[5]), eps=eps,
                         dropout_p=p, seed=seed, x_is_tf32=x_clean, round_out=0, out=None, saved=_p(saved),
                         saved_floats=saved.numel(), ws=_p(ws), ws_floats=n_ws, w1_tf32=_p(ctx.twins[0]),
                         w2_tf32=_p(ctx.twins[1]), dtype=dt, x_h16=_p(ctx.x16), out_h16=None)
        mixed = dt == _lib.DTYPE_F32_H16
        dx_amax = torch.empty(1, device=xc.device, dtype=torch.float32) if mixed else None    # cleared by the operator
        a = _lib.FfnBwdArgs(f=f, dout=_p(dout), dx=_p(dx), dw1=_p(grads[0]), db1=_p(grads[1]), dw2=_p(grads[2]),
                            db2=_p(grads[3]), dln_g=_p(grads[4]), dln_b=_p(grads[5]), grads_zeroed=zeroed,
                            dout_amax=_p(_tagged(dout, _AMAX_ATTR)) if mixed else None, dx_amax=_p(dx_amax))
        check(lib.st_ffn_bwd(C.byref(a), _stream()))
        if mixed:
            _tag(dx, _AMAX_ATTR, dx_amax)
        _notify(ctx.sinks, direct)
        if direct is not None:
            grads = [None] * len(params)
        return (dx, *grads, None, None, None, None)


def positionwise_ffn(x, w1, b1, w2, b2, ln_g, ln_b, eps: float = 1e-6, dropout_p: float = 0.0, seed: int = 0,
                     round_out: bool = True, return_hidden: bool = False):
    """PositionwiseFeedForward.forward (SubLayers.py:24-28) as one fused operator.

    return_hidden=True additionally returns the (non-differentiable) hidden activation
    dropout1(relu(fc1(x))) that backward will use — a test hook: parity of the gradient of a
    piecewise-linear function is only defined for a given ReLU gate pattern."""
    out, hidden = _PositionwiseFFN.apply(x, w1, b1, w2, b2, ln_g, ln_b, eps, dropout_p, seed, round_out)
    return (out, hidden) if return_hidden else out


# ------------------------------------------------------------------------------------------------
# (D) label-smoothed / soft-target cross entropy
# ------------------------------------------------------------------------------------------------
class _LabelSmoothingCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, one_hot, weight, confidence, padding_idx, size_average):
        logits = _need(logits, "output")
        if logits.dim() != 2:
            raise AssertionError("inputs.dim() == 2")            # Loss.py:51
        N, V = logits.shape
        if not (logits.stride(1) == 1 and (N <= 1 or logits.stride(0) >= V)):   # row-padded views are read in place
            logits = logits.contiguous()
        lib = _lib_for(logits)
        ldl = logits.stride(0) if N > 1 else max(V, 1)
        ldg = (V + 3) // 4 * 4
        target = _contig(_need(target, "target", torch.int64))
        one_hot = _contig(_need(one_hot, "one_hot")).view(-1)
        weight = _contig(_need(weight, "weight")).view(-1)
        row_loss = torch.empty(max(N, 1), device=logits.device, dtype=torch.float32)
        loss = torch.empty((), device=logits.device, dtype=torch.float32)
        # gradient rows padded to a multiple of 4 floats: a consumer GEMM (st_linear_bwd) can read it in place
        grad = torch.zeros(N, ldg, device=logits.device, dtype=torch.float32) if ldg != V else \
            torch.empty(N, V, device=logits.device, dtype=torch.float32)
        check(lib.st_lsce_fwd_bwd(_p(logits), ldl, _p(target), _p(one_hot), _p(weight), float(confidence),
                                  int(padding_idx), int(bool(size_average)), N, V, _p(row_loss), _p(loss), _p(grad), ldg,
                                  _stream()))
        ctx.save_for_backward(grad)
        ctx.V = V
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        return (grad * dloss)[:, :ctx.V], None, None, None, None, None, None


def label_smoothing_ce(logits, target, one_hot, weight, confidence: float, padding_idx: int, size_average: bool = True):
    return _LabelSmoothingCE.apply(logits, target, one_hot, weight, confidence, padding_idx, size_average)


class _SoftTargetCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, q, weight, size_average):
        logits = _contig(_need(logits, "inputs"))
        q = _contig(_need(q, "target"))
        lib = _lib_for(logits)
        N, V = logits.shape
        weight = _contig(_need(weight, "weight")).view(-1)
        row_loss = torch.empty(max(N, 1), device=logits.device, dtype=torch.float32)
        loss = torch.empty((), device=logits.device, dtype=torch.float32)
        grad = torch.empty_like(logits)
        check(lib.st_softce_fwd_bwd(_p(logits), V, _p(q), _p(weight), int(bool(size_average)), N, V, _p(row_loss),
                                    _p(loss), _p(grad), V, _stream()))
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        return grad * dloss, None, None, None


def soft_target_ce(logits, q, weight, size_average: bool = True):
    return _SoftTargetCE.apply(logits, q, weight, size_average)


# ------------------------------------------------------------------------------------------------
# callers either side of the path: encoder front-end, vocabulary projection, embedding
# ------------------------------------------------------------------------------------------------
class _Frontend(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, ln_g, ln_b, pe, eps, dropout_p, seed, round_out, out_dtype):
        ctx.set_materialize_grads(False)   # the auxiliary outputs never carry a gradient: no zero tensors for them
        xc = _contig(_need(x, "inputs"))
        if xc.dim() != 3:
            raise RuntimeError("frontend: inputs must be (batch, frames, feature_dim)")
        if out_dtype not in ACT_DTYPES:
            raise RuntimeError(f"frontend: out_dtype must be float32, float16 or bfloat16, got {out_dtype}")
        dt = ACT_DTYPES[out_dtype]
        lib = _lib_for(xc)
        B, T, k = xc.shape
        params = [_contig(_need(t, "parameter")) for t in (w, b, ln_g, ln_b)]
        d = params[0].shape[0]
        if params[0].shape != (d, k):
            raise RuntimeError(f"frontend: weight {tuple(params[0].shape)} does not match feature_dim {k}")
        if pe is not None:
            pe = _contig(_need(pe, "pe"))
            if pe.shape[-1] != d or pe.numel() // d < T:
                raise RuntimeError(f"frontend: positional table {tuple(pe.shape)} too small for T={T}, d={d}")
        rows = B * T
        n_saved = lib.st_frontend_saved_floats(rows, k, d)
        saved = torch.empty(n_saved, device=xc.device, dtype=torch.float32)
        out = torch.empty(B, T, d, device=xc.device, dtype=out_dtype)
        a = _lib.FrontendArgs(rows=rows, T=T, in_dim=k, d_model=d, x=_p(xc), w=_p(params[0]), b=_p(params[1]),
                              ln_g=_p(params[2]), ln_b=_p(params[3]), pe=_p(pe), eps=float(eps),
                              dropout_p=float(dropout_p), seed=int(seed), round_out=int(round_out), out=_p(out),
                              saved=_p(saved), saved_floats=n_saved, ws=None, ws_floats=0, dtype=dt)
        check(lib.st_frontend_fwd(C.byref(a), _stream()))
        ctx.save_for_backward(xc, saved, pe, *params)
        ctx.sinks = _sinks_of((w, b, ln_g, ln_b), ctx)
        ctx.cfg = (rows, T, k, d, float(eps), float(dropout_p), int(seed), dt)
        if round_out and dt == _lib.DTYPE_F32:
            mark_tf32_clean(out)
        off = lib.st_frontend_hidden_offset(rows, k, d)
        hidden = saved[off:off + rows * d].view(B, T, d)
        ctx.mark_non_differentiable(hidden)
        return out, hidden

    @staticmethod
    def backward(ctx, dout, _dhidden):
        if dout is None:      # only the auxiliary output was used downstream: no gradient flows
            return (None,) * 11
        xc, saved, pe, *params = ctx.saved_tensors
        rows, T, k, d, eps, p, seed, dt = ctx.cfg
        lib = _lib_for(xc)
        dout = _contig(dout)
        if ACT_DTYPES.get(dout.dtype) != dt:
            raise RuntimeError(f"frontend backward: gradient dtype {dout.dtype} does not match the forward output")
        n_ws = lib.st_frontend_ws_floats(rows, k, d)
        ws = torch.empty(n_ws, device=xc.device, dtype=torch.float32)
        dx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        direct, zeroed = _claim(ctx.sinks)
        grads = direct if direct is not None else [torch.empty_like(t) for t in params]
        f = _lib.FrontendArgs(rows=rows, T=T, in_dim=k, d_model=d, x=_p(xc), w=_p(params[0]), b=_p(params[1]),
                              ln_g=_p(params[2]), ln_b=_p(params[3]), pe=_p(pe), eps=eps, dropout_p=p, seed=seed,
                              round_out=0, out=None, saved=_p(saved), saved_floats=saved.numel(), ws=_p(ws),
                              ws_floats=n_ws, dtype=dt)
        a = _lib.FrontendBwdArgs(f=f, dout=_p(dout), dx=_p(dx), dw=_p(grads[0]), db=_p(grads[1]), dln_g=_p(grads[2]),
                                 dln_b=_p(grads[3]), grads_zeroed=zeroed)
        check(lib.st_frontend_bwd(C.byref(a), _stream()))
        _notify(ctx.sinks, direct)
        if direct is not None:
            grads = [None] * 4
        return (dx, *grads, None, None, None, None, None, None)


def frontend(x, w, b, ln_g, ln_b, pe=None, eps: float = 1e-6, dropout_p: float = 0.0, seed: int = 0,
             round_out: bool = None, return_hidden: bool = False, out_dtype=torch.float32):
    """Encoder input front-end, Models.py:28-33,42-44: LayerNorm(Dropout(ReLU(Linear(x)))) + pe[:T] as one operator.
    return_hidden additionally returns Dropout(ReLU(Linear(x))) (non-differentiable test hook, cf. positionwise_ffn).
    out_dtype: element type of the result (the activation type of the layers that follow); x stays fp32."""
    out, hidden = _Frontend.apply(x, w, b, ln_g, ln_b, pe, eps, dropout_p, seed, ROUND_OUT if round_out is None else round_out,
                                  out_dtype)
    return (out, hidden) if return_hidden else out


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        xc = _contig(_need_act(x, "input"))
        dt = ACT_DTYPES[xc.dtype]
        x_clean = int(dt != _lib.DTYPE_F32 or is_tf32_clean(x))
        lib = _lib_for(xc)
        k = xc.shape[-1]
        rows = xc.numel() // k
        wc = _contig(_need(w, "weight"))
        bc = None if b is None else _contig(_need(b, "bias"))
        n = wc.shape[0]
        if wc.shape != (n, k):
            raise RuntimeError(f"linear: weight {tuple(wc.shape)} does not match input features {k}")
        ldy = (n + 3) // 4 * 4
        n_saved = lib.st_linear_saved_floats_dt(dt, rows, k, n, x_clean)
        saved = torch.empty(n_saved, device=xc.device, dtype=torch.float32)
        y = torch.empty(rows, ldy, device=xc.device, dtype=torch.float32)
        a = _lib.LinearArgs(rows=rows, in_dim=k, out_dim=n, x=_p(xc), x_is_tf32=x_clean, w=_p(wc), b=_p(bc), y=_p(y),
                            ldy=ldy, saved=_p(saved), saved_floats=n_saved, ws=None, ws_floats=0, dtype=dt)
        check(lib.st_linear_fwd(C.byref(a), _stream()))
        ctx.save_for_backward(xc, saved, wc, bc)
        ctx.sinks = _sinks_of((w, b), ctx) if b is not None else _sinks_of((w,), ctx)
        ctx.cfg = (rows, k, n, x_clean, dt)
        return y[:, :n].view(*xc.shape[:-1], n)        # a row-padded view when n % 4 != 0 (V = 4337)

    @staticmethod
    def backward(ctx, dy):
        xc, saved, wc, bc = ctx.saved_tensors
        rows, k, n, x_clean, dt = ctx.cfg
        lib = _lib_for(xc)
        dy2 = _need(dy, "grad_output").reshape(rows, n)
        if dy2.stride(1) != 1 or (rows > 1 and dy2.stride(0) < n):
            dy2 = dy2.contiguous()
        lddy = dy2.stride(0) if rows > 1 else n
        n_ws = lib.st_linear_ws_floats_dt(dt, rows, k, n)
        ws = torch.empty(n_ws, device=xc.device, dtype=torch.float32)
        dx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        direct, zeroed = _claim(ctx.sinks)
        if direct is not None:
            dw, db = direct[0], (direct[1] if bc is not None else None)
        else:
            dw = torch.empty_like(wc) if ctx.needs_input_grad[1] else None
            db = torch.empty_like(bc) if (bc is not None and ctx.needs_input_grad[2]) else None
        f = _lib.LinearArgs(rows=rows, in_dim=k, out_dim=n, x=_p(xc), x_is_tf32=x_clean, w=_p(wc), b=_p(bc), y=None, ldy=n,
                            saved=_p(saved), saved_floats=saved.numel(), ws=_p(ws), ws_floats=n_ws, dtype=dt)
        a = _lib.LinearBwdArgs(f=f, dy=_p(dy2), lddy=lddy, dx=_p(dx), dw=_p(dw), db=_p(db), grads_zeroed=zeroed)
        check(lib.st_linear_bwd(C.byref(a), _stream()))
        _notify(ctx.sinks, direct)
        if direct is not None:
            dw = db = None
        return dx, dw, db


def linear(x, weight, bias=None):
    """y = x @ weight.T + bias (nn.Linear, e.g. tgt_word_proj Models.py:145,151) on the tcgen05 GEMM, with autograd.
    x may be fp32 (TF32 operands) or fp16 / bf16; y (the logits) is always fp32."""
    return _Linear.apply(x, weight, bias)


class _Embedding(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, table, pe, padding_idx, round_out, out_dtype):
        if out_dtype not in ACT_DTYPES:
            raise RuntimeError(f"embedding: out_dtype must be float32, float16 or bfloat16, got {out_dtype}")
        dt = ACT_DTYPES[out_dtype]
        idx = _contig(_need(idx, "indices", torch.int64))
        table = _contig(_need(table, "embedding weight"))
        lib = _lib_for(table)
        vocab, d = table.shape
        if idx.dim() != 2:
            raise RuntimeError("embedding: indices must be (batch, length)")
        B, L = idx.shape
        if pe is not None:
            pe = _contig(_need(pe, "pe"))
            if pe.shape[-1] != d or pe.numel() // d < L:
                raise RuntimeError(f"embedding: positional table {tuple(pe.shape)} too small for L={L}, d={d}")
        out = torch.empty(B, L, d, device=table.device, dtype=out_dtype)
        check(lib.st_embed_fwd(_p(idx), _p(table), _p(pe), L, _p(out), B * L, d, vocab, int(round_out), dt, _stream()))
        ctx.save_for_backward(idx)
        ctx.sinks = _sinks_of((table,), ctx)
        ctx.cfg = (vocab, d, int(padding_idx), dt)
        if round_out and dt == _lib.DTYPE_F32:
            mark_tf32_clean(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        vocab, d, padding_idx, dt = ctx.cfg
        dout = _contig(dout)
        if ACT_DTYPES.get(dout.dtype) != dt:
            raise RuntimeError(f"embedding backward: gradient dtype {dout.dtype} does not match the forward output")
        lib = _lib_for(dout)
        direct, zeroed = _claim(ctx.sinks)
        dtable = direct[0] if direct is not None else torch.empty(vocab, d, device=dout.device, dtype=torch.float32)
        check(lib.st_embed_bwd(_p(idx), _p(dout), _p(dtable), idx.numel(), d, vocab, padding_idx, 0 if zeroed else 1, dt,
                               _stream()))
        _notify(ctx.sinks, direct)
        return None, (None if direct is not None else dtable), None, None, None, None


def embedding(idx, table, pe=None, padding_idx: int = -1, round_out: bool = None, out_dtype=torch.float32):
    """table[idx] + pe[:L] (Models.py:84-87); the padding row receives no gradient (nn.Embedding padding_idx).
    out_dtype: element type of the result (the activation type of the layers that follow); the table stays fp32."""
    return _Embedding.apply(idx, table, pe, padding_idx, ROUND_OUT if round_out is None else round_out, out_dtype)


# ------------------------------------------------------------------------------------------------
# CTC head of the joint CTC / attention objective
# ------------------------------------------------------------------------------------------------
class _CTCLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, input_lengths, target_lengths, blank):
        logits = _need(logits, "logits")
        if logits.dim() != 3:
            raise RuntimeError("ctc_loss: logits must be (batch, frames, vocab)")
        B, T, V = logits.shape
        if not (logits.stride(2) == 1 and logits.stride(0) == T * logits.stride(1) and logits.stride(1) >= V):
            logits = logits.contiguous()           # row-padded projections (functional.linear) are read in place
        lib = _lib_for(logits)
        targets = _contig(_need(targets, "targets", torch.int64))
        input_lengths = _contig(_need(input_lengths, "input_lengths", torch.int64))
        target_lengths = _contig(_need(target_lengths, "target_lengths", torch.int64))
        if targets.dim() != 2 or targets.shape[0] != B or input_lengths.shape != (B,) or target_lengths.shape != (B,):
            raise RuntimeError("ctc_loss: targets must be (batch, max_target_length), lengths (batch,)")
        L_max = targets.shape[1]
        n_ws = lib.st_ctc_ws_floats(B, T, L_max)
        ws = torch.empty(n_ws, device=logits.device, dtype=torch.float32)
        nll = torch.empty(B, device=logits.device, dtype=torch.float32)
        check(lib.st_ctc_fwd_bwd(_p(logits), logits.stride(1), _p(targets), max(L_max, 1), _p(input_lengths), _p(target_lengths),
                                 int(blank), B, T, V, L_max, _p(nll), None, None, 0, _p(ws), n_ws, _stream()))
        ctx.save_for_backward(logits, targets, input_lengths, target_lengths, nll, ws)
        ctx.cfg = (B, T, V, L_max, int(blank))
        return nll

    @staticmethod
    def backward(ctx, dnll):
        logits, targets, input_lengths, target_lengths, nll, ws = ctx.saved_tensors
        B, T, V, L_max, blank = ctx.cfg
        lib = _lib_for(logits)
        ldg = (V + 3) // 4 * 4
        grad = torch.empty(B, T, ldg, device=logits.device, dtype=torch.float32)
        scale = _contig(dnll.to(torch.float32))
        check(lib.st_ctc_grad(_p(logits), logits.stride(1), _p(targets), max(L_max, 1), _p(input_lengths), _p(target_lengths),
                              blank, B, T, V, L_max, _p(nll), _p(scale), _p(grad), ldg, _p(ws), ws.numel(), _stream()))
        return grad[:, :, :V], None, None, None, None


def ctc_loss(logits, targets, input_lengths, target_lengths, blank: int = 0, reduction: str = "mean"):
    """CTC negative log-likelihood of `targets` (B, L_max; no blanks) given frame logits (B, T, V); log-softmax inside.
    Reductions as torch.nn.functional.ctc_loss: 'none' (B,), 'sum', 'mean' (each loss divided by its target length,
    then the batch mean).  Utterances without a feasible alignment give +inf and a zero gradient."""
    nll = _CTCLoss.apply(logits, targets, input_lengths, target_lengths, blank)
    if reduction == "none":
        return nll
    if reduction == "sum":
        return nll.sum()
    if reduction == "mean":
        return (nll / target_lengths.clamp(min=1).to(nll.dtype)).mean()
    raise ValueError("reduction must be 'none', 'sum' or 'mean'")
