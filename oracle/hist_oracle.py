"""CPU oracle for the validation confusion matrix.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this; the product path never does.

Restates with numpy the reference's mIoU helpers (arman-rn/RTDA-SemanticSegmentation):

* ``confusion_counts``     <- ``fast_hist``      utils.py:62-89
* ``iou_per_class``        <- ``per_class_iou``  utils.py:92-116
* ``validation_confusion`` <- the per-batch slice validation.py:98-109
* ``mean_iou``             <- validation.py:135-136

Integer arithmetic throughout: parity with the CUDA path is bit-exact.
Pinned against the imported reference by ``oracle/pin_against_reference.py``
and the fixtures in ``tests/golden/``.
"""
from __future__ import annotations

import numpy as np

__all__ = ["confusion_counts", "iou_per_class", "validation_confusion", "mean_iou"]


def confusion_counts(label_true: np.ndarray, label_pred: np.ndarray, n_class: int) -> np.ndarray:
    """n_class x n_class int64 matrix, rows = truth, columns = prediction; any
    pair with either value outside [0, n_class) is dropped (utils.py:78-87)."""
    t = np.asarray(label_true)
    p = np.asarray(label_pred)
    ok = (t >= 0) & (t < n_class) & (p >= 0) & (p < n_class)
    flat = n_class * t[ok].astype(np.int64) + p[ok].astype(np.int64)
    return np.bincount(flat, minlength=n_class * n_class).reshape(n_class, n_class).astype(np.int64)


def validation_confusion(labels: np.ndarray, preds: np.ndarray, n_class: int, ignore_index: int = 255):
    """What one validation batch adds to the matrix (validation.py:98-109):
    drop pixels whose label is ``ignore_index``, flatten, ``fast_hist``."""
    labels = np.asarray(labels)
    preds = np.asarray(preds)
    keep = labels != ignore_index
    return confusion_counts(labels[keep].ravel(), preds[keep].ravel(), n_class)


def iou_per_class(hist: np.ndarray) -> np.ndarray:
    """diag / (row sum + column sum - diag + 1e-5), NaN -> 0 (utils.py:92-116)."""
    hist = np.asarray(hist)
    d = np.diag(hist)
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = d / (hist.sum(axis=1) + hist.sum(axis=0) - d + 1e-5)
    return np.nan_to_num(iou, nan=0.0)


def mean_iou(hist: np.ndarray) -> float:
    """``float(np.nanmean(per_class_iou(hist)))`` over all classes (validation.py:135-136)."""
    return float(np.nanmean(iou_per_class(hist)))
