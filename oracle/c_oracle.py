"""ctypes binding of oracle/lovasz_oracle.c.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this.  The C file restates the reference algorithm with the
documented stable tie order (see its header); this wrapper builds it on demand
with ``make -C oracle`` and exposes numpy-level entry points.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblovasz_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lovasz_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.lovasz_oracle_c.restype = ctypes.c_int
        L.lovasz_oracle_c.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64,
            ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_float,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ]
        L.confusion_oracle_c.restype = ctypes.c_int
        L.confusion_oracle_c.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
        ]
        _lib = L
    return _lib


def lovasz(probas, labels, ignore=255, per_image=False, grad_out=1.0):
    """probas [B,C,H,W] float32, labels [B,H,W] int64 (numpy or torch CPU).
    Returns dict(loss float32, grad [B,C,H,W], class_loss [G,C], present [G,C], n_present [G])."""
    p = np.ascontiguousarray(np.asarray(probas, dtype=np.float32))
    l = np.ascontiguousarray(np.asarray(labels, dtype=np.int64))
    b, c, h, w = p.shape
    groups = b if per_image else 1
    loss = np.zeros(1, np.float32)
    grad = np.zeros_like(p)
    class_loss = np.zeros((groups, c), np.float32)
    present = np.zeros((groups, c), np.int32)
    n_present = np.zeros(groups, np.int32)
    rc = lib().lovasz_oracle_c(
        p.ctypes.data, l.ctypes.data, b, c, h * w, 0 if ignore is None else int(ignore), int(ignore is not None),
        int(per_image), float(grad_out), loss.ctypes.data, grad.ctypes.data, class_loss.ctypes.data,
        present.ctypes.data, n_present.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"lovasz_oracle_c failed: {rc}")
    return dict(loss=loss[0], grad=grad, class_loss=class_loss, present=present, n_present=n_present)


def confusion(labels, preds, n_class=19, ignore=255):
    t = np.ascontiguousarray(np.asarray(labels, dtype=np.int64)).ravel()
    q = np.ascontiguousarray(np.asarray(preds, dtype=np.int64)).ravel()
    hist = np.zeros((n_class, n_class), np.int64)
    rc = lib().confusion_oracle_c(t.ctypes.data, q.ctypes.data, t.size, n_class,
                                  0 if ignore is None else int(ignore), int(ignore is not None), hist.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"confusion_oracle_c failed: {rc}")
    return hist
