"""Import the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  The GPU box has no /root/reference: nothing that
runs there (``-m gpu`` tests, ``smoke()``, ``bench.py``) may call this.  It is
used by ``oracle/pin_against_reference.py`` and ``tests/golden/make_golden.py``
to pin the oracle and to write the committed golden fixtures.

``utils.py`` of the reference imports fvcore and albumentations, which are not
installed; empty stub modules satisfy the imports without touching the two
functions on the path (``fast_hist`` utils.py:62-89, ``per_class_iou`` :92-116).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "losses", "lovasz_loss.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


def load():
    """Returns (lovasz_module, utils_module) of the reference."""
    if not available():
        raise RuntimeError("reference not present at /root/reference")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _stub("fvcore")
    _stub("fvcore.nn", FlopCountAnalysis=object, flop_count_table=lambda *a, **k: "")
    _stub("albumentations", Compose=object)
    _stub("albumentations.pytorch", ToTensorV2=object)
    # the repo has its own top-level names `utils` / `losses`; make sure the
    # reference's are the ones imported here.
    for name in ("utils", "losses", "losses.lovasz_loss", "config"):
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith(REFERENCE_ROOT):
            del sys.modules[name]
    lov = importlib.import_module("losses.lovasz_loss")
    utl = importlib.import_module("utils")
    return lov, utl
