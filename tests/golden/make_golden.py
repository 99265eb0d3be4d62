"""Writes tests/golden/*.npz from the UNMODIFIED reference in /root/reference.

Run in the build container (the only place the reference exists):

    python tests/golden/make_golden.py

For every case the fixture holds the inputs and what the reference returned:
``LovaszSoftmax(ignore)(probas, labels)`` loss and autograd gradient
(lovasz_loss.py:185-200), the per-image composition of SURVEY.md 0.2 where
asked, and ``fast_hist`` / ``per_class_iou`` (utils.py:62-116) for the
confusion-matrix cases.  The fixtures are small (a few hundred KB in total) and
are what the oracle and the CUDA path are checked against on machines that do
not have the reference.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import _reference_import  # noqa: E402


def _softmax_case(seed, b, c, h, w, ignore_frac, n_labels=None, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn((b, c, h, w), generator=g) * scale
    labels = torch.randint(0, n_labels or c, (b, h, w), generator=g)
    if ignore_frac > 0:
        labels[torch.rand((b, h, w), generator=g) < ignore_frac] = 255
    return torch.softmax(logits, 1), labels


def lovasz_cases():
    """name -> (probas fp32 [B,C,H,W], labels int64 [B,H,W], ignore, per_image)."""
    cases = {}
    # KAT1 of SURVEY.md section 4 (hand-checkable)
    ch1 = torch.tensor([0.9, 0.6, 0.3, 0.2])
    p = torch.stack([1 - ch1, ch1]).view(1, 2, 1, 4)
    cases["kat1"] = (p, torch.tensor([[[1, 0, 1, 255]]]), 255, False)
    # KAT2 of SURVEY.md section 4
    torch.manual_seed(1234)
    logits = torch.randn(2, 19, 32, 64)
    labels = torch.randint(0, 19, (2, 32, 64))
    labels[torch.rand(2, 32, 64) < 0.1] = 255
    cases["kat2_global"] = (torch.softmax(logits, 1), labels, 255, False)
    cases["kat2_per_image"] = (torch.softmax(logits, 1), labels, 255, True)
    # odd, non-multiple-of-4 shapes
    cases["odd_3x5x7x13"] = (*_softmax_case(11, 3, 5, 7, 13, 0.15), 255, False)
    cases["odd_3x5x7x13_per_image"] = (*_softmax_case(11, 3, 5, 7, 13, 0.15), 255, True)
    cases["odd_1x19x33x47"] = (*_softmax_case(12, 1, 19, 33, 47, 0.1, scale=3.0), 255, False)
    # no ignore index at all
    cases["no_ignore_2x4x8x8"] = (*_softmax_case(13, 2, 4, 8, 8, 0.0), None, False)
    # ignore index set but no pixel ignored
    cases["none_ignored_2x6x9x10"] = (*_softmax_case(14, 2, 6, 9, 10, 0.0), 255, False)
    # classes absent from the labels (labels only use 3 of 19 classes)
    cases["absent_classes"] = (*_softmax_case(15, 2, 19, 16, 16, 0.1, n_labels=3), 255, False)
    # single pixel (the p == 1 branch of lovasz_grad, lovasz_loss.py:68)
    cases["single_pixel"] = (torch.tensor([0.2, 0.5, 0.3]).view(1, 3, 1, 1), torch.tensor([[[1]]]), 255, False)
    # a class with exactly one pixel among many
    pr, lb = _softmax_case(16, 1, 4, 8, 8, 0.0, n_labels=3)
    lb[0, 3, 3] = 3
    cases["one_pixel_class"] = (pr, lb, 255, False)
    # everything ignored -> the reference returns the int 0
    pr, lb = _softmax_case(17, 2, 5, 4, 4, 0.0)
    cases["all_ignored"] = (pr, torch.full_like(lb, 255), 255, False)
    # one image fully ignored in per-image mode
    pr, lb = _softmax_case(18, 3, 5, 6, 6, 0.1)
    lb[1] = 255
    cases["per_image_one_empty"] = (pr, lb, 255, True)
    # larger, confident (trained-like) predictions
    g = torch.Generator().manual_seed(19)
    lb = torch.randint(0, 19, (2, 8, 8), generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2)
    lb[torch.rand((2, 64, 64), generator=g) < 0.05] = 255
    lg = torch.randn((2, 19, 64, 64), generator=g) * 3
    lg.scatter_add_(1, lb.clamp(max=18).unsqueeze(1), torch.full((2, 1, 64, 64), 6.0))
    cases["confident_2x19x64x64"] = (torch.softmax(lg, 1), lb, 255, False)
    return cases


def tie_cases():
    """Inputs with massive ties: the reference's gradient inside a tie group is
    implementation-defined (unstable sort), so these fixtures pin the LOSS only."""
    cases = {}
    g = torch.Generator().manual_seed(21)
    lb = torch.randint(0, 5, (2, 12, 12), generator=g)
    lb[torch.rand((2, 12, 12), generator=g) < 0.1] = 255
    cases["uniform_probs"] = (torch.full((2, 5, 12, 12), 0.2), lb, 255, False)
    onehot = torch.zeros(2, 5, 12, 12)
    pred = torch.randint(0, 5, (2, 12, 12), generator=g)
    onehot.scatter_(1, pred.unsqueeze(1), 1.0)
    cases["onehot_probs"] = (onehot, lb, 255, False)
    q = torch.randint(0, 4, (2, 5, 12, 12), generator=g).float() / 4.0
    cases["quantised_probs"] = (q, lb, 255, False)
    return cases


def midsize_cases():
    """Mid-size cases regenerated from the synthetic generators at test time (the inputs are not stored:
    10 MB each); the fixture holds what the reference returned for them: the loss, the sum of |gradient|
    per class, 4096 sampled gradient entries, and a checksum of the inputs so that a drift of the
    generator is noticed.  name -> (b, h, w, seed, scale, boost, per_image)."""
    return {
        "mid_conf_1x19x256x512": (1, 256, 512, 7001, 3.0, 6.0, False),           # D_conf: > 1 bucket per segment, k_plan's linear maps
        "mid_saturated_1x19x192x384": (1, 192, 384, 7002, 3.0, 16.0, False),     # boost 16: exact ties at 1 - k * 2^-24
        "mid_conf_per_image_2x19x160x256": (2, 160, 256, 7003, 3.0, 6.0, True),
    }


def midsize_inputs(b, h, w, seed, scale, boost):
    from rtda_semanticsegmentation_b200 import synthetic

    labels = synthetic.blocky_labels(b, h, w, seed=seed)
    logits = synthetic.confident_logits(labels, seed=seed, scale=scale, boost=boost)
    return torch.softmax(logits, 1), labels


def hist_cases():
    cases = {}
    g = np.random.default_rng(31)
    t = g.integers(0, 19, size=5000)
    p = g.integers(0, 19, size=5000)
    cases["uniform_19"] = (t, p, 19)
    t2 = t.copy()
    t2[g.random(5000) < 0.1] = 255
    p2 = p.copy()
    p2[g.random(5000) < 0.05] = 19  # out of range prediction
    p2[g.random(5000) < 0.05] = -1  # negative prediction
    t2[g.random(5000) < 0.02] = -3
    cases["out_of_range_19"] = (t2, p2, 19)
    cases["diagonal_7"] = (np.repeat(np.arange(7), 11), np.repeat(np.arange(7), 11), 7)
    cases["empty_19"] = (np.zeros(0, np.int64), np.zeros(0, np.int64), 19)
    cases["two_class"] = (g.integers(0, 2, size=333), g.integers(0, 2, size=333), 2)
    return cases


def main():
    lov, utl = _reference_import.load()
    torch.set_num_threads(1)  # fixed reduction order for the recorded values

    def run_ref(p, l, ignore, per_image):
        crit = lov.LovaszSoftmax(ignore=ignore)
        x = p.clone().requires_grad_(True)
        if per_image:
            loss = lov.mean(crit(x[i:i + 1], l[i:i + 1]) for i in range(x.size(0)))
        else:
            loss = crit(x, l)
        if torch.is_tensor(loss) and loss.requires_grad:
            loss.backward()
            return float(loss), np.float32(loss.item()), x.grad.numpy().copy(), True
        return 0.0, np.float32(0), np.zeros_like(p.numpy()), False

    out = {}
    meta = []
    for name, (p, l, ignore, per_image) in {**lovasz_cases(), **tie_cases()}.items():
        p = p.float().contiguous()
        l = l.long().contiguous()
        _, loss32, grad, is_tensor = run_ref(p, l, ignore, per_image)
        out[f"{name}/probas"] = p.numpy()
        out[f"{name}/labels"] = l.numpy()
        out[f"{name}/loss"] = np.asarray(loss32, np.float32)
        out[f"{name}/grad"] = grad.astype(np.float32)
        out[f"{name}/ignore"] = np.asarray(-1 if ignore is None else ignore, np.int64)
        out[f"{name}/per_image"] = np.asarray(int(per_image), np.int64)
        out[f"{name}/ref_returned_tensor"] = np.asarray(int(is_tensor), np.int64)
        out[f"{name}/ties"] = np.asarray(int(name in tie_cases()), np.int64)
        meta.append(name)
    # lovasz_grad known answers (lovasz_loss.py:47-70)
    for i, gt in enumerate([[1, 0, 1, 0, 0], [1], [0, 0, 1], [1, 1, 1], [0, 1, 0, 1, 1, 0, 0, 1]]):
        out[f"lovasz_grad/{i}/gt"] = np.asarray(gt, np.float32)
        out[f"lovasz_grad/{i}/out"] = lov.lovasz_grad(torch.tensor(gt, dtype=torch.float32)).numpy()
    out["lovasz_grad/count"] = np.asarray(5, np.int64)
    np.savez_compressed(os.path.join(HERE, "lovasz_golden.npz"), names=np.asarray(meta), **out)

    mout = {}
    mnames = []
    for name, (b, h, w, seed, scale, boost, per_image) in midsize_cases().items():
        p, l = midsize_inputs(b, h, w, seed, scale, boost)
        _, loss32, grad, _ = run_ref(p, l, 255, per_image)
        rng = np.random.default_rng(seed)
        nz = np.flatnonzero(grad.ravel())
        idx = np.sort(np.concatenate([rng.choice(nz, 3072, replace=False), rng.integers(0, grad.size, 1024)]))
        mout[f"{name}/spec"] = np.asarray([b, h, w, seed, per_image], np.int64)
        mout[f"{name}/scale_boost"] = np.asarray([scale, boost], np.float64)
        mout[f"{name}/input_checksum"] = np.asarray([int(p.view(torch.int32).to(torch.int64).sum()), int(l.sum())], np.int64)
        mout[f"{name}/loss"] = np.asarray(loss32, np.float32)
        mout[f"{name}/abs_grad_per_class"] = np.abs(grad.astype(np.float64)).sum(axis=(0, 2, 3))
        mout[f"{name}/grad_idx"] = idx.astype(np.int64)
        mout[f"{name}/grad_val"] = grad.ravel()[idx].astype(np.float32)
        mout[f"{name}/grad_max"] = np.asarray(np.abs(grad).max(), np.float32)
        mout[f"{name}/exact_one_probabilities"] = np.asarray(int((p == 1.0).sum()), np.int64)
        mnames.append(name)
    np.savez_compressed(os.path.join(HERE, "lovasz_midsize_golden.npz"), names=np.asarray(mnames), **mout)

    hout = {}
    hnames = []
    for name, (t, p, n) in hist_cases().items():
        t = np.asarray(t, np.int64)
        p = np.asarray(p, np.int64)
        h = utl.fast_hist(t, p, n)
        hout[f"{name}/true"] = t
        hout[f"{name}/pred"] = p
        hout[f"{name}/n_class"] = np.asarray(n, np.int64)
        hout[f"{name}/hist"] = np.asarray(h, np.int64)
        hout[f"{name}/iou"] = np.asarray(utl.per_class_iou(h.astype(np.float64)), np.float64)
        hnames.append(name)
    # the validation.py:98-109 slice on a 2-D map with ignore pixels
    g = np.random.default_rng(32)
    lab = g.integers(0, 19, size=(2, 24, 40))
    lab[g.random(lab.shape) < 0.1] = 255
    prd = g.integers(0, 19, size=lab.shape)
    valid = lab != 255
    h = utl.fast_hist(lab[valid].flatten(), prd[valid].flatten(), 19)
    hout["val_slice/labels"] = lab.astype(np.int64)
    hout["val_slice/preds"] = prd.astype(np.int64)
    hout["val_slice/hist"] = np.asarray(h, np.int64)
    hout["val_slice/iou"] = np.asarray(utl.per_class_iou(h.astype(np.float64)), np.float64)
    hout["val_slice/miou"] = np.asarray(float(np.nanmean(utl.per_class_iou(h.astype(np.float64)))), np.float64)
    np.savez_compressed(os.path.join(HERE, "hist_golden.npz"), names=np.asarray(hnames), **hout)
    print("wrote", len(meta), "lovasz cases,", len(mnames), "mid-size cases and", len(hnames) + 1, "hist cases")


if __name__ == "__main__":
    main()
