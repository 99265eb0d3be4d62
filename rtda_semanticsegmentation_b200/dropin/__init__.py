"""Drop-in seam for the reference's import paths.

The reference's callers do ``from losses.lovasz_loss import LovaszSoftmax`` (main.py:27,
main_adversarial.py:20) and ``from utils import fast_hist, per_class_iou`` (validation.py:19).
Two ways to swap the B200 path in without touching them:

1. put this directory FIRST on ``sys.path`` (``PYTHONPATH=.../rtda_semanticsegmentation_b200/dropin``):
   ``losses/lovasz_loss.py`` here shadows the reference's module, same names, same signatures;
2. call ``install()`` after the reference's modules are imported: it rebinds
   ``losses.lovasz_loss.LovaszSoftmax`` (+ ``lovasz_grad``, ``mean``, ``isnan``) and
   ``utils.fast_hist`` / ``utils.per_class_iou`` (and the names already imported into
   ``validation``) to the CUDA-backed implementations.
"""
from __future__ import annotations

import sys

from .. import lovasz_loss as _ll
from .. import metrics as _m

_LOSS_NAMES = ("LovaszSoftmax", "lovasz_grad", "mean", "isnan")
_METRIC_NAMES = ("fast_hist", "per_class_iou")


def install(loss=True, metrics=True):
    """Rebind the reference's already-imported modules to the B200 implementations.
    Returns the list of ``module.name`` bindings that were replaced."""
    done = []
    if loss:
        mod = sys.modules.get("losses.lovasz_loss")
        if mod is None:
            from .losses import lovasz_loss as mod  # our shim under the reference's name
            pkg = sys.modules.setdefault("losses", sys.modules[__name__ + ".losses"])
            setattr(pkg, "lovasz_loss", mod)
            sys.modules["losses.lovasz_loss"] = mod
            done.append("losses.lovasz_loss (module)")
        for name in _LOSS_NAMES:
            setattr(mod, name, getattr(_ll, name))
            done.append(f"losses.lovasz_loss.{name}")
        for caller in ("main", "main_adversarial", "__main__"):
            m = sys.modules.get(caller)
            if m is not None and getattr(m, "LovaszSoftmax", None) is not None and m.LovaszSoftmax is not _ll.LovaszSoftmax:
                m.LovaszSoftmax = _ll.LovaszSoftmax
                done.append(f"{caller}.LovaszSoftmax")
    if metrics:
        for modname in ("utils", "validation"):
            m = sys.modules.get(modname)
            if m is None:
                continue
            for name in _METRIC_NAMES:
                if hasattr(m, name):
                    setattr(m, name, getattr(_m, name))
                    done.append(f"{modname}.{name}")
    return done
