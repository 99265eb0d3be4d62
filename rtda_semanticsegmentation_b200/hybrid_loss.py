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
    def forward(ctx, logits, labels, ignore, per_image, holder, shard_group=None, global_images=None,
                ce_reduction="global", grad_reduction="sum"):
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
        ctx.lov_scale = ctx.ce_scale = None
        loss_ce, loss_lov, n_valid = outs.loss_ce, outs.loss, outs.n_valid
        if shard_group is not None:
            # this rank's images are a slice of a batch sharded over the ranks (dist.py): ONE all-reduce(sum)
            # of the packed partials {sum of image losses, images, class sums x 2C, CE sum, valid pixels}
            import torch.distributed as dist

            from .lovasz_loss import _pending_reduce

            grp = None if shard_group is True else shard_group
            world = dist.get_world_size(grp)
            packed = outs._view("packed")
            k = 2 + 2 * c
            dscale = float(world) if grad_reduction == "mean" else 1.0   # DDP averages gradients over the ranks
            if ce_reduction == "global":
                # exact single-device semantics: CE = sum of all ranks' -log p / number of all ranks' valid
                # pixels.  Its gradient needs the global pixel count, so the collective is in stream order.
                from .dist import losses_from_packed

                dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=grp)
                loss_lov, loss_ce, _ = losses_from_packed(packed, c)
                ctx.lov_scale = (float(b) * dscale / packed[1]).float()
                ctx.ce_scale = dscale if dscale != 1.0 else None
                n_valid = packed[k + 1].round().to(torch.int32).reshape(1)
            else:
                # "rank_mean": CE = mean over ranks of each rank's own mean (what DDP computes with the
                # reference).  Nothing on this rank's data path depends on the collective: it runs
                # asynchronously and both returned losses are this rank's SHARE of the global value.
                if global_images is None:
                    raise ValueError("ce_reduction='rank_mean' needs global_images (the batch size over all ranks)")
                loss_lov = (packed[0].clone() / float(global_images)).float()
                loss_ce = outs.loss_ce / float(world)
                ctx.lov_scale = float(b) * dscale / float(global_images)
                ctx.ce_scale = dscale / float(world)
                prev = _pending_reduce.pop(dev.index, None)
                if prev is not None and prev._work is not None:
                    prev._packed()
                outs._work = dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=grp, async_op=True)
                _pending_reduce[dev.index] = outs
        ctx.save_for_backward(probas, unit, codes, outs.class_present, outs.n_present, n_valid)
        return loss_ce, loss_lov

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
        if ctx.lov_scale is not None:
            glv = glv * ctx.lov_scale
        if ctx.ce_scale is not None:
            gce = gce * ctx.ce_scale
        grad = torch.empty((b, c, h, w), dtype=torch.float32, device=dev)
        if torch.cuda.current_device() != dev.index:
            torch.cuda.set_device(dev)
        _lib.check(lib.rtda_hybrid_bwd(
            gce.data_ptr(), glv.data_ptr(), probas.data_ptr(), unit.data_ptr(), codes.data_ptr(),
            present.data_ptr(), n_present.data_ptr(), n_valid.data_ptr(), b, c, h * w, int(ctx.per_image),
            grad.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "rtda_hybrid_bwd")
        return grad, None, None, None, None, None, None, None, None


def hybrid_ce_lovasz(logits, labels, ignore=255, lovasz_weight=0.5, per_image=False, return_outputs=False,
                     shard_group=None, global_images=None, ce_reduction="global", grad_reduction="sum"):
    """Returns (loss_total, loss_ce, loss_lovasz): ``CE(logits, labels, ignore_index=ignore) +
    lovasz_weight * LovaszSoftmax(ignore)(softmax(logits, 1), labels)`` as train_lovasz.py:98-105
    computes them; all three are 0-dim fp32 CUDA tensors, ``loss_total`` carries the graph.

    ``shard_group`` (a process group, or True for the default one; needs ``per_image=True``): ``logits`` /
    ``labels`` are this rank's slice of a batch sharded over the ranks (BASELINE configs[2] and [4]).  One
    all-reduce(sum) of 4+2C doubles; the backward needs none.  ``ce_reduction``:

    * ``"global"`` (default) - exactly the single-device numbers: the cross-entropy is averaged over the valid
      pixels of ALL ranks, the Lovasz term over all images; every rank gets the global losses.  The
      collective is in stream order (the CE gradient needs the global pixel count).
    * ``"rank_mean"`` (needs ``global_images``) - the cross-entropy is the mean over ranks of each rank's own
      mean, which is what DDP computes with the reference.  The collective is asynchronous and the
      returned losses are this rank's SHARE of the global values (the shares sum to them).

    ``grad_reduction``: ``"sum"`` (default) scales the local gradient so that the SUM over ranks of the
    parameter gradients is the single-device gradient; ``"mean"`` multiplies it by the world size, for
    DistributedDataParallel, which averages gradients."""
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
    if shard_group is not None and not per_image:
        raise ValueError("only per_image=True shards over ranks (global mode needs one sort over all pixels)")
    if ce_reduction not in ("global", "rank_mean") or grad_reduction not in ("sum", "mean"):
        raise ValueError("ce_reduction is 'global' or 'rank_mean', grad_reduction 'sum' or 'mean'")
    if (logits.data_ptr() % 16 and (logits.size(2) * logits.size(3)) % 4 == 0):
        logits = logits.clone()          # a view at an odd offset: the float4 paths need 16-byte alignment
    holder = []
    loss_ce, loss_lovasz = _HybridFunction.apply(logits, labels, ignore, bool(per_image), holder, shard_group,
                                                 global_images, ce_reduction, grad_reduction)
    total = loss_ce + lovasz_weight * loss_lovasz
    if return_outputs:
        return total, loss_ce, loss_lovasz, holder[0]
    return total, loss_ce, loss_lovasz


class HybridCELovasz(nn.Module):
    """``criterion(logits, labels) -> loss_total`` with ``last_ce`` / ``last_lovasz`` kept for logging
    (train_lovasz.py:133-135 reads the three losses with ``.item()``).  ``shard_group`` /
    ``global_images`` / ``ce_reduction`` / ``grad_reduction``: see :func:`hybrid_ce_lovasz`."""

    def __init__(self, ignore=255, lovasz_weight=0.5, per_image=False, shard_group=None, global_images=None,
                 ce_reduction="global", grad_reduction="sum", capture=False):
        super().__init__()
        # capture=True: one CUDA graph of the forward and one of the backward per input shape (graphed.py);
        # not with shard_group (the collective is not captured)
        if capture and shard_group is not None:
            raise ValueError("capture=True does not cover the sharded form (its all-reduce is not captured)")
        self._graphed = None
        if capture:
            from .graphed import GraphedCriterion

            self._graphed = GraphedCriterion(lambda x, l: hybrid_ce_lovasz(x, l, self.ignore, self.lovasz_weight, self.per_image))
        self.ignore = ignore
        self.lovasz_weight = lovasz_weight
        self.per_image = per_image
        self.shard_group = shard_group
        self.global_images = global_images
        self.ce_reduction = ce_reduction
        self.grad_reduction = grad_reduction
        self.last_ce = None
        self.last_lovasz = None
        self.last_outputs = None

    def forward(self, logits, labels):
        if self._graphed is not None and not torch.cuda.is_current_stream_capturing():
            if logits.dtype != torch.float32:
                logits = logits.float()
            if labels.dtype not in (torch.int64, torch.int32, torch.uint8):
                labels = labels.long()
            total, ce, lov = self._graphed(logits.contiguous(), labels)
            self.last_ce, self.last_lovasz, self.last_outputs = ce.detach(), lov.detach(), None
            return total
        total, ce, lov, (outs, _probas) = hybrid_ce_lovasz(
            logits, labels, self.ignore, self.lovasz_weight, self.per_image, return_outputs=True,
            shard_group=self.shard_group, global_images=self.global_images, ce_reduction=self.ce_reduction,
            grad_reduction=self.grad_reduction)
        self.last_ce, self.last_lovasz, self.last_outputs = ce.detach(), lov.detach(), outs
        return total
