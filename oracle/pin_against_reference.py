"""Pins the oracle against the UNMODIFIED reference (build container only).

    python oracle/pin_against_reference.py

TEST INFRASTRUCTURE ONLY.  Differential check of ``oracle/lovasz_oracle.py``
and ``oracle/hist_oracle.py`` against ``/root/reference`` on seeded random
inputs: loss and gradient of ``LovaszSoftmax.forward`` (lovasz_loss.py:185-200),
``lovasz_grad`` (:47-70) bit-for-bit incl. the integer restatement, ``mean``
(:19-44), ``fast_hist`` / ``per_class_iou`` (utils.py:62-116) bit-for-bit.
Exits non-zero on any mismatch.  The same comparison runs as a CPU test
(tests/test_oracle_pin.py) whenever /root/reference is present.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import _reference_import, hist_oracle, lovasz_oracle  # noqa: E402


def check_all(verbose=True):
    lov, utl = _reference_import.load()
    torch.set_num_threads(1)
    failures = []

    def note(ok, what):
        if verbose:
            print(("ok   " if ok else "FAIL ") + what)
        if not ok:
            failures.append(what)

    # lovasz_grad: fp32 restatement and integer restatement, bit for bit
    rng = np.random.default_rng(0)
    for n in [1, 2, 3, 17, 256, 5000, 100000]:
        for density in [0.02, 0.5, 0.97]:
            gt = (rng.random(n) < density).astype(np.float32)
            if gt.sum() == 0:
                gt[rng.integers(0, n)] = 1
            ref = lov.lovasz_grad(torch.from_numpy(gt)).numpy()
            a = lovasz_oracle.jaccard_grad(torch.from_numpy(gt)).numpy()
            b = lovasz_oracle.jaccard_grad_int(gt)
            note(np.array_equal(ref.view(np.uint32), a.view(np.uint32)), f"jaccard_grad fp32 n={n} d={density}")
            note(np.array_equal(ref.view(np.uint32), b.view(np.uint32)), f"jaccard_grad_int n={n} d={density}")

    # mean
    def vals():  # fresh tensors each time: the reference's `acc += v` mutates the first one in place
        return [torch.tensor(v) for v in (0.3, 0.55, 0.125)]
    note(float(lov.mean(iter(vals()))) == float(lovasz_oracle.seq_mean(iter(vals()))), "seq_mean 3 values")
    note(lov.mean(iter([])) == lovasz_oracle.seq_mean(iter([])), "seq_mean empty")
    note(float(lov.mean(iter(vals()[:1]))) == float(lovasz_oracle.seq_mean(iter(vals()[:1]))), "seq_mean single")

    # loss + gradient, unstable order like the reference (same torch, same ops -> bitwise)
    shapes = [(1, 19, 16, 24, 0.1), (2, 19, 32, 64, 0.1), (3, 5, 7, 13, 0.2), (2, 4, 8, 8, 0.0), (4, 19, 24, 24, 0.3)]
    for seed, (b, c, h, w, ign) in enumerate(shapes):
        g = torch.Generator().manual_seed(100 + seed)
        p = torch.softmax(torch.randn((b, c, h, w), generator=g) * 2, 1)
        l = torch.randint(0, c, (b, h, w), generator=g)
        if ign:
            l[torch.rand((b, h, w), generator=g) < ign] = 255
        for per_image in (False, True):
            x = p.clone().requires_grad_(True)
            crit = lov.LovaszSoftmax(ignore=255)
            if per_image:
                ref = lov.mean(crit(x[i:i + 1], l[i:i + 1]) for i in range(b))
            else:
                ref = crit(x, l)
            ref.backward()
            for stable in (False, True):
                loss, grad = lovasz_oracle.lovasz_softmax_with_grad(p, l, ignore=255, per_image=per_image, stable=stable)
                tag = f"lovasz {b}x{c}x{h}x{w} per_image={per_image} stable={stable}"
                note(abs(float(loss) - float(ref)) <= 1e-6 * abs(float(ref)), tag + " loss")
                # random softmax outputs: ties are rare, stable and unstable agree to rounding
                note(float((grad - x.grad).abs().max()) <= 1e-6 * float(x.grad.abs().max()), tag + " grad")
                if not stable:
                    note(torch.equal(grad, x.grad), tag + " grad bitwise")

    # confusion matrix
    for seed in range(4):
        r = np.random.default_rng(200 + seed)
        n = [19, 19, 7, 2][seed]
        t = r.integers(-2, n + 2, size=4000)
        q = r.integers(-2, n + 2, size=4000)
        t[r.random(4000) < 0.1] = 255
        ref = utl.fast_hist(t, q, n)
        got = hist_oracle.confusion_counts(t, q, n)
        note(np.array_equal(ref, got), f"fast_hist n={n}")
        note(np.array_equal(utl.per_class_iou(ref.astype(np.float64)), hist_oracle.iou_per_class(got.astype(np.float64))),
             f"per_class_iou n={n}")
    return failures


if __name__ == "__main__":
    bad = check_all()
    print("failures:", len(bad))
    sys.exit(1 if bad else 0)
