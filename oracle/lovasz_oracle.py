"""CPU oracle for the Lovasz-Softmax hot path.  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path
(``rtda_semanticsegmentation_b200``) never imports anything from ``oracle/``
and raises if its CUDA library is missing.

It restates, on CPU with plain torch / numpy ops, the algorithm of the
reference file ``losses/lovasz_loss.py`` (arman-rn/RTDA-SemanticSegmentation):

* ``seq_mean``              <- ``mean``                 lovasz_loss.py:19-44
* ``jaccard_grad``          <- ``lovasz_grad``          lovasz_loss.py:47-70
* ``jaccard_grad_int``      integer restatement of the same (bit-identical
                            for segments shorter than 2**24, see below)
* ``flatten_nchw``          <- ``prob_flatten``         lovasz_loss.py:93-120
* ``lovasz_flat``           <- ``lovasz_softmax_flat``  lovasz_loss.py:122-183
* ``lovasz_softmax``        <- ``LovaszSoftmax.forward`` lovasz_loss.py:185-200
                            plus the per-image composition SURVEY.md 0.2
                            defines (the reference has no per_image mode).

Parity status: the reference ships no tests and no golden vectors, so this
oracle is pinned (a) against the *unmodified reference imported from
/root/reference* by ``oracle/pin_against_reference.py`` (run in the build
container; the GPU box has no /root/reference) and (b) against the fixtures
that script's sibling ``tests/golden/make_golden.py`` wrote from the reference.

Tie order.  ``torch.sort(descending=True)`` on CPU is not stable, so the
reference's per-pixel gradient inside a group of *equal* errors is
implementation-defined.  ``stable=True`` here pins the order the CUDA path
documents: descending error, ties by ascending flat pixel index.  The loss is
tie-invariant up to rounding; only gradients inside mixed fg/non-fg tie groups
depend on it.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = [
    "seq_mean",
    "jaccard_grad",
    "jaccard_grad_int",
    "flatten_nchw",
    "class_terms",
    "lovasz_flat",
    "lovasz_softmax",
    "lovasz_softmax_with_grad",
]


def seq_mean(values, empty=0):
    """Sequential mean, the way the reference's ``mean`` does it
    (lovasz_loss.py:19-44): first value, then ``acc += v`` in order, then one
    division by the count; a single value is returned undivided; no values
    returns ``empty`` (the reference's Python int 0)."""
    acc = None
    n = 0
    for v in values:
        n += 1
        acc = v if acc is None else acc + v
    if n == 0:
        return empty
    if n == 1:
        return acc
    return acc / n


def jaccard_grad(fg_sorted: torch.Tensor) -> torch.Tensor:
    """Gradient of the Lovasz extension of the Jaccard loss w.r.t. sorted
    errors (lovasz_loss.py:47-70): J_i = 1 - (G - cumsum fg)/(G + cumsum(1-fg)),
    g_0 = J_0, g_i = J_i - J_{i-1}; fp32 throughout like the reference."""
    n = fg_sorted.numel()
    total = fg_sorted.sum()
    inter = total - torch.cumsum(fg_sorted, 0)
    union = total + torch.cumsum(1.0 - fg_sorted, 0)
    jac = 1.0 - inter / union
    if n > 1:  # lovasz_loss.py:68 single-element guard
        jac = torch.cat([jac[:1], jac[1:] - jac[:-1]])
    return jac


def jaccard_grad_int(fg_sorted: np.ndarray) -> np.ndarray:
    """Integer restatement of ``jaccard_grad``: the cumulative sums are exact
    integers, converted to fp32 once, followed by one IEEE fp32 division and
    two fp32 subtractions.  Bit-identical to the fp32 original whenever the
    segment is shorter than 2**24 (fp32 holds those integers exactly); this is
    the formulation the CUDA kernel evaluates per element from
    (rank, fg-count-so-far)."""
    fg = np.asarray(fg_sorted).astype(np.int64)
    n = fg.size
    if n == 0:
        return np.zeros(0, np.float32)
    total = int(fg.sum())
    f_incl = np.cumsum(fg)
    rank1 = np.arange(1, n + 1, dtype=np.int64)
    inter = (total - f_incl).astype(np.float32)
    union = (total + rank1 - f_incl).astype(np.float32)
    jac = (np.float32(1.0) - inter / union).astype(np.float32)
    out = jac.copy()
    if n > 1:
        out[1:] = jac[1:] - jac[:-1]
    return out


def flatten_nchw(probas: torch.Tensor, labels: torch.Tensor):
    """[B,C,H,W] -> [B*H*W, C] and [B,H,W] -> [B*H*W] (lovasz_loss.py:93-120)."""
    if probas.dim() != 4:
        raise ValueError(f"Expected 4D input, but got {probas.dim()}D")
    c = probas.size(1)
    return probas.permute(0, 2, 3, 1).contiguous().view(-1, c), labels.view(-1)


def class_terms(probas: torch.Tensor, labels: torch.Tensor, ignore=None, stable=True):
    """Per-class intermediates of ``lovasz_flat`` for present classes, in
    ascending class order: list of dicts with the class id, the pixel indices
    of the valid pixels (into the flat input), the errors, the sort permutation
    and the Jaccard gradient vector.  Follows lovasz_loss.py:142-180."""
    out = []
    if probas.numel() == 0:
        return out
    present = torch.unique(labels).tolist()  # :142
    if ignore is not None:
        keep = labels != ignore  # :157
        keep_idx = torch.nonzero(keep).view(-1)
    else:
        keep = None
        keep_idx = torch.arange(labels.numel())
    for c in present:
        if ignore is not None and c == ignore:  # :150-153
            continue
        fg = (labels == c).float()  # :147
        if keep is not None:
            p_c = probas[keep, c]  # :160
            fg = fg[keep]  # :161
        else:
            p_c = probas[:, c]
        if fg.sum() == 0:  # :167-168
            continue
        err = (fg - p_c).abs()  # :171
        if stable:
            err_sorted, perm = torch.sort(err, dim=0, descending=True, stable=True)
        else:
            err_sorted, perm = torch.sort(err, 0, descending=True)  # :173, the reference's call
        fg_sorted = fg[perm]  # :176
        grad = jaccard_grad(fg_sorted)
        out.append(
            dict(cls=int(c), pix=keep_idx, err=err, err_sorted=err_sorted, perm=perm,
                 fg_sorted=fg_sorted, grad=grad, loss=torch.dot(err_sorted, grad))  # :180
        )
    return out


def lovasz_flat(probas: torch.Tensor, labels: torch.Tensor, ignore=None, stable=True):
    """Flat Lovasz-Softmax over [P,C] probabilities and [P] labels, classes
    'present' (lovasz_loss.py:122-183).  Empty input returns ``probas * 0``
    (:135-136); no present class returns the reference's int 0."""
    if probas.numel() == 0:
        return probas * 0.0
    return seq_mean(t["loss"] for t in class_terms(probas, labels, ignore, stable))


def lovasz_softmax(probas, labels, ignore=None, per_image=False, stable=True):
    """``LovaszSoftmax(ignore=ignore)(probas, labels)`` (lovasz_loss.py:185-200).
    ``per_image=True`` is the composition SURVEY.md 0.2 defines:
    ``mean(forward(p[i:i+1], l[i:i+1]) for i in range(B))`` with the
    reference's own ``mean``."""
    if per_image:
        return seq_mean(
            lovasz_flat(*flatten_nchw(probas[i:i + 1], labels[i:i + 1]), ignore=ignore, stable=stable)
            for i in range(probas.size(0))
        )
    return lovasz_flat(*flatten_nchw(probas, labels), ignore=ignore, stable=stable)


def lovasz_softmax_with_grad(probas, labels, ignore=None, per_image=False, stable=True,
                             grad_out=1.0):
    """Loss and d(loss * grad_out)/d(probas) through torch autograd on CPU.
    Returns (loss: 0-dim fp32 tensor, grad: tensor like probas).  A degenerate
    input (nothing valid) yields a zero loss and zero gradient."""
    p = probas.detach().clone().float().requires_grad_(True)
    loss = lovasz_softmax(p, labels, ignore=ignore, per_image=per_image, stable=stable)
    if not torch.is_tensor(loss) or loss.dim() != 0 or not loss.requires_grad:
        return torch.zeros((), dtype=torch.float32), torch.zeros_like(p)
    (loss * grad_out).backward()
    return loss.detach(), p.grad.detach()
