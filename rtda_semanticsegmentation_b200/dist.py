"""Multi-GPU composition of the hot path: one process per GPU, torch.distributed (NCCL).

The reference is single-device (config.py:103) and has no collective.  The path
shards only where it is naturally partitioned:

* per-image Lovasz: images are independent sort problems, so each rank runs the
  kernels on its slice of the batch and ONE all-reduce(sum) of a small fp64
  buffer ``[sum of image losses, image count, per-class loss sums x C,
  per-class presence counts x C]`` gives the global mean.  The backward needs no
  collective: d(loss)/d(loss_b) = 1/B_global on the rank that owns image b.
* confusion matrix: counts are additive, one all-reduce(sum) of n*n int64.
* global-mode Lovasz does NOT shard (it needs one sort per class over all
  pixels): under data parallelism every rank computes the loss of its local
  batch, which is what DDP does with the reference ("replicas only").

The functions below take tensors on whatever device the process group works on
(CUDA/NCCL in production, CPU/gloo in the unit tests of the packing logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["pack_partials", "ShardedMeanOfImages", "sharded_per_image_mean", "lovasz_softmax_sharded",
           "hybrid_ce_lovasz_sharded", "losses_from_packed",
           "all_reduce_confusion", "shard_range"]


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced slice [lo, hi) of ``n_items`` for ``rank`` of ``world``."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_partials(group_loss: torch.Tensor, class_loss: torch.Tensor, class_present: torch.Tensor,
                  ce_sum=None, n_valid=None) -> torch.Tensor:
    """fp64 buffer [sum_b loss_b, B_local, sum_b class_loss[b,:], sum_b present[b,:], CE sum, valid pixels]
    - the layout k_finalize writes (include/rtda_lovasz.h, `packed`); the last two are 0 without the
    cross-entropy term."""
    c = class_loss.size(-1)
    buf = torch.zeros(4 + 2 * c, dtype=torch.float64, device=group_loss.device)
    buf[0] = group_loss.double().sum()
    buf[1] = float(group_loss.numel())
    buf[2:2 + c] = class_loss.double().view(-1, c).sum(0)
    buf[2 + c:2 + 2 * c] = class_present.double().view(-1, c).sum(0)
    if ce_sum is not None:
        buf[2 + 2 * c] = float(ce_sum)
        buf[3 + 2 * c] = float(n_valid)
    return buf


def losses_from_packed(packed: torch.Tensor, n_class: int):
    """(Lovasz loss, cross-entropy, valid pixels) of the WHOLE batch from the all-reduced partial buffer:
    mean over all images of the per-image losses; sum of -log p_label over all valid pixels / their number."""
    k = 2 + 2 * n_class
    return (packed[0] / packed[1]).float(), (packed[k] / packed[k + 1]).float(), packed[k + 1]


class ShardedMeanOfImages(torch.autograd.Function):
    """mean over ALL ranks' images of the per-image losses; one all-reduce forward,
    none backward."""

    @staticmethod
    def forward(ctx, local_sum, packed, group):
        # packed[0] is the detached value of local_sum; the collective sums it over ranks
        red = packed.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(red, op=dist.ReduceOp.SUM, group=group)
        ctx.save_for_backward(red)
        ctx.mark_non_differentiable(red)
        return (red[0] / red[1]).to(local_sum.dtype), red

    @staticmethod
    def backward(ctx, grad_loss, _grad_packed):
        (packed,) = ctx.saved_tensors
        return (grad_loss.double() / packed[1]).to(grad_loss.dtype), None, None


def sharded_per_image_mean(group_loss, class_loss, class_present, group=None):
    """Global per-image mean from this rank's per-image losses ``group_loss`` [B_local]
    (differentiable w.r.t. ``group_loss``).  Returns (loss, stats) where stats holds the
    reduced per-class loss means and presence counts for logging."""
    packed = pack_partials(group_loss.detach(), class_loss.detach(), class_present.detach())
    loss, packed = ShardedMeanOfImages.apply(group_loss.sum(), packed, group)
    c = class_loss.size(-1)
    counts = packed[2 + c:2 + 2 * c]
    stats = dict(images=packed[1], class_loss_mean=packed[2:2 + c] / counts.clamp(min=1), class_count=counts)
    return loss, stats


def lovasz_softmax_sharded(probas_local, labels_local, ignore=255, group=None, global_images=None,
                           grad_reduction="sum"):
    """per_image=True Lovasz-Softmax over a batch sharded across ranks: each rank passes its own
    images; every rank gets the global loss; gradients stay local.  The finalize kernel writes the
    partial buffer ``[sum of image losses, image count, class loss sums, class presence counts]``
    and the autograd Function all-reduces it once (csrc/lovasz.cu k_finalize, lovasz_loss.py).

    ``global_images`` (the batch size summed over the ranks, when the caller knows it - a fixed
    per-rank batch): the gradient then does not depend on the collective at all, so the all-reduce is
    issued asynchronously and the ranks no longer meet at a barrier inside every step.  The returned
    loss is this rank's share ``sum(local image losses) / global_images`` (the shares sum to the global
    loss; gradients are identical) - NOT the loss itself; ``stats["loss"]`` is the global loss and waits
    for the collective.

    Gradient convention: the local gradient is scaled by ``1 / B_global``, which is right when the ranks'
    parameter gradients are SUMMED.  ``DistributedDataParallel`` AVERAGES them; pass
    ``grad_reduction="mean"`` there (the local gradient is then multiplied by the world size), otherwise
    the parameter gradients come out ``world_size`` times smaller than on one device.

    Memory: the sort workspace (0.55 GB for 4x19x512x1024, 3.9 GB for 16x19x720x1280) is cached per
    (device, stream) until ``lovasz_loss.release_workspaces()``."""
    from .lovasz_loss import lovasz_softmax

    loss, outs = lovasz_softmax(probas_local, labels_local, per_image=True, ignore=ignore, return_outputs=True,
                                shard_group=True if group is None else group, global_images=global_images,
                                grad_reduction=grad_reduction)
    return loss, _ShardStats(outs, probas_local.size(1))


class _ShardStats(dict):
    """Lazily evaluated logging statistics of a sharded forward (device tensors, no sync)."""

    def __init__(self, outs, c):
        super().__init__()
        self._outs, self._c = outs, c

    def __missing__(self, key):
        packed, c = self._outs.packed, self._c
        counts = packed[2 + c:2 + 2 * c]
        vals = dict(images=packed[1], loss=(packed[0] / packed[1]).float(),
                    class_loss_mean=packed[2:2 + c] / counts.clamp(min=1), class_count=counts,
                    valid_pixels=packed[3 + 2 * c], loss_ce=(packed[2 + 2 * c] / packed[3 + 2 * c].clamp(min=1)).float())
        self.update(vals)
        return vals[key]


def hybrid_ce_lovasz_sharded(logits_local, labels_local, ignore=255, lovasz_weight=0.5, group=None,
                             global_images=None, ce_reduction="global", grad_reduction="sum"):
    """The training step's loss (train_lovasz.py:98-105) over a batch sharded across ranks, per-image
    Lovasz semantics: ``CE + lovasz_weight * Lovasz`` from this rank's logits / labels; ONE all-reduce of
    4+2C doubles; the backward is local.  With ``ce_reduction="global"`` (default) every rank gets
    exactly the single-device numbers (cross-entropy over ALL ranks' valid pixels, Lovasz mean over all
    images); ``"rank_mean"`` (needs ``global_images``) keeps the collective off the data path and
    returns this rank's shares.  See hybrid_loss.hybrid_ce_lovasz.  Returns (total, ce, lovasz, stats)."""
    from .hybrid_loss import hybrid_ce_lovasz

    total, ce, lov, (outs, _probas) = hybrid_ce_lovasz(
        logits_local, labels_local, ignore=ignore, lovasz_weight=lovasz_weight, per_image=True, return_outputs=True,
        shard_group=True if group is None else group, global_images=global_images, ce_reduction=ce_reduction,
        grad_reduction=grad_reduction)
    return total, ce, lov, _ShardStats(outs, logits_local.size(1))


def all_reduce_confusion(hist: torch.Tensor, group=None) -> torch.Tensor:
    """Sum an int64 [n,n] confusion matrix over the ranks in place (one collective)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist
