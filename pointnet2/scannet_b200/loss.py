"""Weighted cross-entropy of the semseg training step -- mirror of
/root/reference/lib/loss.py:5-18 (``WeightedCrossEntropyLoss``: per-point
``F.cross_entropy(reduction="none", ignore_index)`` times per-point weights, then the mean).

On CUDA fp32 logits with ``num_classes % 4 == 0`` the loss and its adjoint are two kernels of
libpn2_b200.so (csrc/loss.cu: one read of the logits forward; one read + one write backward);
anything else -- CPU tensors (the benchmark's cpu_baseline leg), other dtypes, odd class counts --
takes the reference's own torch formulation.  ``fused.ENABLED = False`` forces the latter (the
benchmark's reference arm must not run this repository's kernels).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import fused
from . import lib as _lib

_counters = {}


def _counter(dev):
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    buf = _counters.get(key)
    if buf is None:
        buf = _counters[key] = torch.zeros(1, dtype=torch.int32, device=dev)
    return buf


class _WeightedCE(Function):
    @staticmethod
    def forward(ctx, logits, targets, weights, ignore_index):
        L = _lib.load()
        P, C = logits.shape
        dev = logits.device
        lse = torch.empty(P, dtype=torch.float32, device=dev)
        part = torch.empty(L.pn2_weighted_ce_partials(P), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.pn2_weighted_ce_fwd(
                P, C, ignore_index, logits.data_ptr(), targets.data_ptr(),
                weights.data_ptr() if weights is not None else None, lse.data_ptr(), part.data_ptr(),
                _counter(dev).data_ptr(), loss.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                "weighted_ce_fwd")
        ctx.saved = (logits, targets, weights, lse, ignore_index)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        logits, targets, weights, lse, ignore_index = ctx.saved
        L = _lib.load()
        P, C = logits.shape
        dev = logits.device
        g = grad_loss.to(torch.float32).contiguous()
        dlogits = torch.empty_like(logits)
        with torch.cuda.device(dev):
            _lib.check(L.pn2_weighted_ce_bwd(
                P, C, ignore_index, logits.data_ptr(), targets.data_ptr(),
                weights.data_ptr() if weights is not None else None, lse.data_ptr(), g.data_ptr(),
                dlogits.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "weighted_ce_bwd")
        return dlogits, None, None, None


def kernel_eligible(logits, targets, weights):
    return (fused.ENABLED and logits.is_cuda and logits.dim() == 2 and logits.dtype == torch.float32
            and logits.size(1) % 4 == 0 and targets.dtype == torch.int64 and targets.dim() == 1
            and (weights is None or (weights.dtype == torch.float32 and weights.dim() == 1)))


def weighted_cross_entropy(logits, targets, weights=None, ignore_index=-100):
    """logits (P, classes), targets (P,), weights (P,) or None -> scalar (lib/loss.py:10-18)."""
    if kernel_eligible(logits, targets, weights):
        return _WeightedCE.apply(logits.contiguous(), targets.contiguous(),
                                 weights.contiguous() if weights is not None else None, int(ignore_index))
    loss = F.cross_entropy(logits, targets, reduction="none", ignore_index=ignore_index)
    return torch.mean(loss * weights) if weights is not None else torch.mean(loss)


class WeightedCrossEntropyLoss(nn.Module):
    """Same constructor and call signature as the reference's (lib/loss.py:5-18)."""

    def __init__(self, ignore_index=-100):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, inputs, targets, weights=None):
        assert inputs.size(0) == targets.size(0) and (weights is None or weights.size(0) == inputs.size(0))
        return weighted_cross_entropy(inputs, targets, weights, self.ignore_index)
