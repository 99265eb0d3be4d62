"""Drop-in for the reference's mIoU helpers on B200.

* ``fast_hist(label_true, label_pred, n_class)``  <- utils.py:62-89 (same signature;
  numpy arrays or tensors in, numpy int64 [n,n] out);
* ``per_class_iou(hist)``                         <- utils.py:92-116 (host, fp64);
* ``confusion_matrix(labels, preds, n_class, ignore_index)`` - the device-resident
  form of the validation.py:98-109 slice: CUDA int64 maps in, CUDA int64 [n,n] out,
  no host copy of the maps;
* ``confusion_matrix_from_logits`` - the same fused with the argmax of validation.py:95;
* ``ConfusionMatrix`` - the accumulator a validation loop keeps (validation.py:70,105).

The counting runs in ``csrc/confusion.cu`` through ``rtda_confusion_accumulate``;
there is no CPU path (numpy inputs are copied to the current CUDA device).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

__all__ = ["fast_hist", "per_class_iou", "confusion_matrix", "confusion_matrix_from_logits",
           "ConfusionMatrix", "mean_iou"]


_INT_KINDS = {torch.int64: _lib.LABEL_I64, torch.uint8: _lib.LABEL_U8, torch.int32: _lib.LABEL_I32}


def _device(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("rtda_semanticsegmentation_b200 needs a CUDA device: there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _as_device_ints(x, device):
    if isinstance(x, np.ndarray):
        if x.dtype not in (np.int64, np.int32, np.uint8):
            x = x.astype(np.int64)
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not torch.is_tensor(x):
        x = torch.as_tensor(x)
    if x.dtype not in (torch.int64, torch.int32, torch.uint8):
        x = x.long()
    if not x.is_cuda:
        x = x.to(device, non_blocking=True)
    return x.contiguous()


def confusion_matrix(labels, preds, n_class=19, ignore_index=255, out=None):
    """hist[t, p] += 1 over pixels with label != ignore_index, 0 <= t,p < n_class
    (validation.py:102-107 + utils.py:78-87).  Returns a CUDA int64 [n,n] tensor;
    pass ``out`` to accumulate into an existing one."""
    lib = _lib.load()
    if out is not None:
        # the kernel does 64-bit atomics on n_class^2 slots of `out`: anything else is an out-of-bounds write
        if (not torch.is_tensor(out) or out.dtype != torch.int64 or not out.is_cuda or not out.is_contiguous()
                or out.numel() != n_class * n_class):
            raise ValueError("out must be a contiguous CUDA int64 tensor with n_class*n_class elements")
    if (out is not None and torch.is_tensor(labels) and torch.is_tensor(preds) and labels.is_cuda and preds.is_cuda
            and labels.dtype in _INT_KINDS and preds.dtype in _INT_KINDS and labels.is_contiguous()
            and preds.is_contiguous() and labels.numel() == preds.numel()
            and labels.device == out.device and preds.device == out.device):
        # fast path of a validation loop: device maps, existing accumulator - one C call
        dev = labels.device
        if torch.cuda.current_device() != dev.index:
            torch.cuda.set_device(dev)
        _lib.check(lib.rtda_confusion_accumulate(
            labels.data_ptr(), _INT_KINDS[labels.dtype], preds.data_ptr(), _INT_KINDS[preds.dtype], labels.numel(), n_class,
            0 if ignore_index is None else int(ignore_index), int(ignore_index is not None), out.data_ptr(),
            torch.cuda.current_stream(dev).cuda_stream), "rtda_confusion_accumulate")
        return out
    dev = labels.device if torch.is_tensor(labels) and labels.is_cuda else _device()
    labels = _as_device_ints(labels, dev)
    preds = _as_device_ints(preds, dev)
    if labels.numel() != preds.numel():
        raise ValueError(f"labels {tuple(labels.shape)} and preds {tuple(preds.shape)} differ in size")
    if torch.is_tensor(preds) and preds.is_cuda and preds.device != dev:
        raise ValueError(f"labels are on {dev} and preds on {preds.device}")
    if out is None:
        out = torch.zeros((n_class, n_class), dtype=torch.int64, device=dev)
    elif out.device != dev:
        raise ValueError(f"out is on {out.device} and the maps on {dev}")
    with torch.cuda.device(dev):
        _lib.check(lib.rtda_confusion_accumulate(
            labels.data_ptr(), _lib.label_kind(labels.dtype), preds.data_ptr(), _lib.label_kind(preds.dtype),
            labels.numel(), n_class, 0 if ignore_index is None else int(ignore_index), int(ignore_index is not None),
            out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "rtda_confusion_accumulate")
    return out


def confusion_matrix_from_logits(logits, labels, ignore_index=255, out=None):
    """Fuses ``preds = argmax(logits, 1)`` (validation.py:95) into the histogram:
    logits [B,n_class,H,W] fp32 CUDA, labels [B,H,W]."""
    lib = _lib.load()
    if logits.dim() != 4:
        raise ValueError(f"Expected 4D logits, but got {logits.dim()}D")
    if not logits.is_cuda:
        raise RuntimeError("logits must be a CUDA tensor: there is no CPU path")
    dev = logits.device
    logits = logits.float().contiguous()
    labels = _as_device_ints(labels, dev)
    b, n_class, h, w = logits.shape
    if labels.numel() != b * h * w:
        raise ValueError("labels do not match logits")
    if out is None:
        out = torch.zeros((n_class, n_class), dtype=torch.int64, device=dev)
    elif (not torch.is_tensor(out) or out.dtype != torch.int64 or not out.is_cuda or not out.is_contiguous()
          or out.numel() != n_class * n_class or out.device != dev):
        raise ValueError("out must be a contiguous CUDA int64 tensor with n_class*n_class elements on the logits' device")
    with torch.cuda.device(dev):
        _lib.check(lib.rtda_confusion_argmax_accumulate(
            logits.data_ptr(), labels.data_ptr(), _lib.label_kind(labels.dtype), b, n_class, h * w,
            0 if ignore_index is None else int(ignore_index), int(ignore_index is not None),
            out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "rtda_confusion_argmax_accumulate")
    return out


def fast_hist(label_true, label_pred, n_class):
    """Reference signature (utils.py:62-89): flattened integer arrays in, numpy
    int64 [n_class, n_class] out.  No ignore index here - like the reference, the
    caller filters first (validation.py:102-107) or relies on the range mask."""
    h = confusion_matrix(label_true, label_pred, n_class=n_class, ignore_index=None)
    return h.cpu().numpy()


def per_class_iou(hist):
    """diag / (row + col - diag + 1e-5), NaN -> 0, on the host in fp64 (utils.py:92-116)."""
    if torch.is_tensor(hist):
        hist = hist.detach().cpu().numpy()
    hist = np.asarray(hist)
    tp = np.diag(hist)
    with np.errstate(divide="ignore", invalid="ignore"):
        ious = tp / (hist.sum(axis=1) + hist.sum(axis=0) - tp + 1e-5)
    return np.nan_to_num(ious, nan=0.0)


def mean_iou(hist) -> float:
    """Mean over ALL classes, absent ones counting 0 (validation.py:135-136)."""
    return float(np.nanmean(per_class_iou(hist)))


class ConfusionMatrix:
    """Device-resident accumulator for a validation loop (validation.py:70,105,135-136).

    ``update(labels, preds)`` / ``update_from_logits(logits, labels)`` enqueue one
    kernel each and never synchronise; ``compute()`` copies the 19x19 counts to the
    host once.  ``all_reduce()`` sums the counts over a process group (one NCCL
    all-reduce of n*n int64) when validation images are sharded over ranks."""

    def __init__(self, n_class=19, ignore_index=255, device=None):
        self.n_class = n_class
        self.ignore_index = ignore_index
        self.hist = torch.zeros((n_class, n_class), dtype=torch.int64, device=_device(device))

    def reset(self):
        self.hist.zero_()

    def update(self, labels, preds):
        confusion_matrix(labels, preds, self.n_class, self.ignore_index, out=self.hist)
        return self

    def update_from_logits(self, logits, labels):
        confusion_matrix_from_logits(logits, labels, self.ignore_index, out=self.hist)
        return self

    def all_reduce(self, group=None):
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.hist, op=dist.ReduceOp.SUM, group=group)
        return self

    def compute(self):
        return self.hist.cpu().numpy()

    def per_class_iou(self):
        return per_class_iou(self.compute().astype(np.float64))

    def mean_iou(self):
        return mean_iou(self.compute().astype(np.float64))
