"""B200-native Lovasz-Softmax loss and confusion-matrix histogram.

A from-scratch sm_100a implementation of the one data-parallel hot path of
arman-rn/RTDA-SemanticSegmentation (losses/lovasz_loss.py and utils.py:62-116),
behind the reference's own call boundary.  See DESIGN.md and INTEGRATION.md.
"""
from . import _lib
from .hybrid_loss import HybridCELovasz, hybrid_ce_lovasz
from .lovasz_loss import (LovaszSoftmax, flatten_probas, isnan, lovasz_grad, lovasz_softmax,
                          lovasz_softmax_flat, mean, segment_sort_paths, set_sort_path)
from .step_metrics import StepMetrics
from .metrics import (ConfusionMatrix, confusion_matrix, confusion_matrix_from_logits, fast_hist,
                      mean_iou, per_class_iou)

__all__ = [
    "LovaszSoftmax", "lovasz_softmax", "lovasz_softmax_flat", "lovasz_grad", "flatten_probas", "mean", "isnan",
    "fast_hist", "per_class_iou", "mean_iou", "confusion_matrix", "confusion_matrix_from_logits",
    "ConfusionMatrix", "HybridCELovasz", "hybrid_ce_lovasz", "build", "native_library_path",
    "set_sort_path", "segment_sort_paths", "StepMetrics",
]


def build(verbose=False):
    """Compile the in-tree CUDA library (nvcc, sm_100a)."""
    return _lib.build(verbose)


def native_library_path():
    return _lib.LIB_PATH
