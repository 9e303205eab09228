"""Shared test utilities: golden fixtures, the parity metric, oracle <-> module parameter mapping."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance for fp32 activations and gradients: max|a-b| / max|b| <= 1e-3
TOL = 1e-3


def golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def t(a, device="cpu", dtype=None):
    x = torch.from_numpy(np.asarray(a))
    if dtype is not None:
        x = x.to(dtype)
    return x.to(device)


def relerr(a, b):
    """max|a-b| / max|b| (the parity metric of SURVEY.md §8c); NaNs must coincide."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    nan_a, nan_b = torch.isnan(a), torch.isnan(b)
    if not torch.equal(nan_a, nan_b):
        return float("inf")
    a, b = a[~nan_a], b[~nan_b]
    if a.numel() == 0:
        return 0.0
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


def relu_gate_from_cuda(hidden_cuda, pre_oracle):
    """Gate pattern (0/1) the CUDA forward used, checked against the oracle's pre-activations.

    The gradient of a piecewise-linear function is only comparable for a fixed set of active ReLU units:
    any reduced-precision forward (TF32 here) may put a pre-activation that lies within its rounding
    error of 0 on the other side of the kink, which flips a whole gradient term.  So: (1) assert that
    the CUDA gate differs from the oracle's only where |pre| is tiny (<= 1% of the pre-activation std)
    and only for a small fraction of units, (2) return the CUDA gate for the oracle to differentiate with.
    """
    gate = hidden_cuda.detach().cpu() > 0
    pre = pre_oracle.detach().cpu()
    assert gate.shape == pre.shape, (gate.shape, pre.shape)
    flips = gate != (pre > 0)
    if flips.any():
        tau = 1e-2 * pre.double().std().item()
        assert pre[flips].abs().max().item() <= tau, ("ReLU gate differs away from the kink", pre[flips].abs().max().item(), tau)
        assert flips.double().mean().item() < 1e-2
    return gate.to(torch.float64)


def params(g, prefix="p."):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def torch_params(g, prefix="p.", device="cpu", dtype=torch.float32, requires_grad=False):
    out = {}
    for k, v in params(g, prefix).items():
        x = t(v, device, dtype)
        out[k] = x.requires_grad_() if requires_grad else x
    return out
