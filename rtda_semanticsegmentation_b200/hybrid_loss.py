"""Fused cross-entropy + Lovasz-Softmax from LOGITS.

Every caller of the reference's loss does the same thing around it
(train_lovasz.py:98-105, :120-127, :280-289):

    loss_ce     = criterion_ce(logits, labels)            # nn.CrossEntropyLoss(ignore_index=255)
    probas      = F.softmax(logits, dim=1)
    loss_lovasz = criterion_lovasz(probas, labels)        # LovaszSoftmax(ignore=255)
    loss_total  = loss_ce + lovasz_weight * loss_lovasz   # config.py:292 LOVASZ_LOSS_WEIGHT = 0.5

``HybridCELovasz`` computes the same three numbers and the same gradient w.r.t. the logits with
the softmax and the cross-entropy sum folded into the Lovasz histogram kernel's read of the logits
(rtda_hybrid_fwd) and softmax-backward + CE-backward folded into one gradient kernel
(rtda_hybrid_bwd): the probabilities make one round trip through HBM instead of four, and the
~8 elementwise torch kernels around the loss disappear.  No CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .lovasz_loss import LovaszOutputs, _require_cuda, _workspace

__all__ = ["HybridCELovasz", "hybrid_ce_lovasz"]


class _HybridFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, ignore, per_image, holder):
        lib = _lib.load()
        b, c, h, w = logits.shape
        hw = h * w
        groups = b if per_image else 1
        dev = logits.device
        if torch.cuda.current_device() != dev.index:
            torch.cuda.set_device(dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        nbytes = _lib.ctypes.c_size_t(0)
        _lib.check(lib.rtda_lovasz_workspace_bytes(b, c, hw, int(per_image), _lib.ctypes.byref(nbytes)),
                   "rtda_lovasz_workspace_bytes")
        ws = _workspace(dev, stream, nbytes.value)
        probas = torch.empty((b, c, h, w), dtype=torch.float32, device=dev)
        unit = torch.empty(b * c * hw, dtype=torch.float32, device=dev)
        codes = torch.empty(b * hw, dtype=torch.uint8, device=dev)
        outs = LovaszOutputs(torch.empty(LovaszOutputs.nbytes(groups, c), dtype=torch.uint8, device=dev), groups, c)
        _lib.check(lib.rtda_hybrid_fwd(
            logits.data_ptr(), labels.data_ptr(), _lib.label_kind(labels.dtype),
            b, c, hw, 0 if ignore is None else int(ignore), int(ignore is not None), int(per_image),
            ws.data_ptr(), nbytes.value, probas.data_ptr(), unit.data_ptr(), codes.data_ptr(),
            outs.ptr("loss"), outs.ptr("loss_ce"), outs.ptr("class_loss"), outs.ptr("class_present"),
            outs.ptr("n_present"), outs.ptr("group_loss"), outs.ptr("label_errors"), outs.ptr("packed"),
            outs.ptr("n_valid"), stream), "rtda_hybrid_fwd")
        if holder is not None:
            holder.append((outs, probas))
        ctx.shape = (b, c, h, w)
        ctx.per_image = per_image
        ctx.save_for_backward(probas, unit, codes, outs.class_present, outs.n_present, outs.n_valid)
        return outs.loss_ce, outs.loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_ce, g_lov):
        probas, unit, codes, present, n_present, n_valid = ctx.saved_tensors
        b, c, h, w = ctx.shape
        lib = _lib.load()
        dev = probas.device
        zero = torch.zeros(1, dtype=torch.float32, device=dev)
        gce = zero if g_ce is None else g_ce.detach().to(device=dev, dtype=torch.float32).reshape(1)
        glv = zero if g_lov is None else g_lov.detach().to(device=dev, dtype=torch.float32).reshape(1)
        grad = torch.empty((b, c, h, w), dtype=torch.float32, device=dev)
        if torch.cuda.current_device() != dev.index:
            torch.cuda.set_device(dev)
        _lib.check(lib.rtda_hybrid_bwd(
            gce.data_ptr(), glv.data_ptr(), probas.data_ptr(), unit.data_ptr(), codes.data_ptr(),
            present.data_ptr(), n_present.data_ptr(), n_valid.data_ptr(), b, c, h * w, int(ctx.per_image),
            grad.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "rtda_hybrid_bwd")
        return grad, None, None, None, None


def hybrid_ce_lovasz(logits, labels, ignore=255, lovasz_weight=0.5, per_image=False, return_outputs=False):
    """Returns (loss_total, loss_ce, loss_lovasz): ``CE(logits, labels, ignore_index=ignore) +
    lovasz_weight * LovaszSoftmax(ignore)(softmax(logits, 1), labels)`` as train_lovasz.py:98-105
    computes them; all three are 0-dim fp32 CUDA tensors, ``loss_total`` carries the graph."""
    if logits.dim() != 4:
        raise ValueError(f"Expected 4D input, but got {logits.dim()}D")
    _require_cuda(logits, "logits")
    _require_cuda(labels, "labels")
    if logits.dtype == torch.float64:
        raise TypeError("logits must be float32 / float16 / bfloat16")
    if logits.dtype != torch.float32:
        logits = logits.float()          # softmax and CE run in fp32 under autocast as well (train_lovasz.py:101)
    logits = logits.contiguous()
    labels.view(-1)
    if labels.dtype not in (torch.int64, torch.int32, torch.uint8):
        labels = labels.long()
    if labels.numel() != logits.size(0) * logits.size(2) * logits.size(3):
        raise ValueError(f"labels {tuple(labels.shape)} do not match logits {tuple(logits.shape)}")
    holder = []
    loss_ce, loss_lovasz = _HybridFunction.apply(logits, labels, ignore, bool(per_image), holder)
    total = loss_ce + lovasz_weight * loss_lovasz
    if return_outputs:
        return total, loss_ce, loss_lovasz, holder[0]
    return total, loss_ce, loss_lovasz


class HybridCELovasz(nn.Module):
    """``criterion(logits, labels) -> loss_total`` with ``last_ce`` / ``last_lovasz`` kept for logging
    (train_lovasz.py:133-135 reads the three losses with ``.item()``)."""

    def __init__(self, ignore=255, lovasz_weight=0.5, per_image=False):
        super().__init__()
        self.ignore = ignore
        self.lovasz_weight = lovasz_weight
        self.per_image = per_image
        self.last_ce = None
        self.last_lovasz = None

    def forward(self, logits, labels):
        total, ce, lov = hybrid_ce_lovasz(logits, labels, self.ignore, self.lovasz_weight, self.per_image)
        self.last_ce, self.last_lovasz = ce.detach(), lov.detach()
        return total
