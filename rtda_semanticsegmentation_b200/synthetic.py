"""Synthetic Cityscapes / GTA5-shaped inputs for tests and benchmarks.

No dataset or checkpoint exists on the GPU box, so every measurement uses the
generators SURVEY.md section 8(d) specifies:

* ``blocky_labels``  - spatially coherent label maps (32x32 tiles of one class,
  10 % of tiles and 2 % of single pixels set to the ignore index), the shape
  real segmentation masks have;
* ``confident_logits`` (D_conf, headline) - ``3*randn + 6*onehot(label)``: a
  trained-like prediction; the exact tail-pruning shortcut keeps ~93 % of the
  elements here, so the benchmark pays for the full sort;
* ``diffuse_case`` (D_diffuse) - iid labels, ``randn`` logits;
* ``validation_maps`` - label map + prediction map that agrees on 80 % of pixels.

Everything is generated with a ``torch.Generator`` on the requested device so
the GPU benchmark never touches the host for its device-resident leg.
"""
from __future__ import annotations

import torch

IGNORE_INDEX = 255  # config.py:62
NUM_CLASSES = 19    # config.py:61


def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def blocky_labels(b, h, w, n_class=NUM_CLASSES, ignore=IGNORE_INDEX, seed=0, device="cpu",
                  tile=32, ignore_tiles=0.10, ignore_pixels=0.02):
    """[b,h,w] int64 label map made of ``tile`` x ``tile`` blocks."""
    g = _gen(device, seed)
    th, tw = -(-h // tile), -(-w // tile)
    coarse = torch.randint(0, n_class, (b, th, tw), generator=g, device=device)
    if ignore is not None and ignore_tiles > 0:
        drop = torch.rand((b, th, tw), generator=g, device=device) < ignore_tiles
        coarse = torch.where(drop, torch.full_like(coarse, ignore), coarse)
    lab = coarse.repeat_interleave(tile, 1).repeat_interleave(tile, 2)[:, :h, :w].contiguous()
    if ignore is not None and ignore_pixels > 0:
        lab[torch.rand((b, h, w), generator=g, device=device) < ignore_pixels] = ignore
    return lab


def confident_logits(labels, n_class=NUM_CLASSES, ignore=IGNORE_INDEX, seed=0, scale=3.0, boost=6.0):
    """D_conf logits [b,n_class,h,w] fp32 for a label map (ignored pixels get a
    uniformly drawn boosted class)."""
    device = labels.device
    g = _gen(device, seed + 7919)
    b, h, w = labels.shape
    logits = torch.randn((b, n_class, h, w), generator=g, device=device) * scale
    tgt = labels.clone()
    bad = (tgt < 0) | (tgt >= n_class)
    rnd = torch.randint(0, n_class, (b, h, w), generator=g, device=device)
    tgt = torch.where(bad, rnd, tgt)
    logits.scatter_add_(1, tgt.unsqueeze(1), torch.full((b, 1, h, w), boost, device=device))
    return logits


def confident_case(b, h, w, n_class=NUM_CLASSES, ignore=IGNORE_INDEX, seed=0, device="cpu"):
    """(probas fp32 [b,C,h,w], labels int64 [b,h,w]) from D_conf."""
    labels = blocky_labels(b, h, w, n_class, ignore, seed, device)
    probas = torch.softmax(confident_logits(labels, n_class, ignore, seed), dim=1)
    return probas, labels


def diffuse_case(b, h, w, n_class=NUM_CLASSES, ignore=IGNORE_INDEX, seed=0, device="cpu", ignore_frac=0.10):
    """(probas, labels) from D_diffuse: iid labels, softmax(randn)."""
    g = _gen(device, seed)
    labels = torch.randint(0, n_class, (b, h, w), generator=g, device=device)
    if ignore is not None and ignore_frac > 0:
        labels[torch.rand((b, h, w), generator=g, device=device) < ignore_frac] = ignore
    probas = torch.softmax(torch.randn((b, n_class, h, w), generator=g, device=device), dim=1)
    return probas, labels


def validation_maps(b, h, w, n_class=NUM_CLASSES, ignore=IGNORE_INDEX, seed=0, device="cpu", wrong=0.20):
    """(labels, preds) int64 [b,h,w]: blocky labels; predictions equal the label
    except on ``wrong`` of the pixels (and on ignored pixels), which get a
    uniformly drawn class - the shape of a real validation confusion matrix
    (mass on the diagonal, which is what stresses the histogram's atomics)."""
    labels = blocky_labels(b, h, w, n_class, ignore, seed, device)
    g = _gen(device, seed + 104729)
    rnd = torch.randint(0, n_class, (b, h, w), generator=g, device=device)
    flip = torch.rand((b, h, w), generator=g, device=device) < wrong
    bad = (labels < 0) | (labels >= n_class)
    preds = torch.where(flip | bad, rnd, labels)
    return labels, preds
