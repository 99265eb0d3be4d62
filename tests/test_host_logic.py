"""CPU: host-side mirror of the reference interface (no compute calls)."""
import numpy as np
import pytest
import torch

import rtda_semanticsegmentation_b200 as rtda
from rtda_semanticsegmentation_b200 import synthetic
from rtda_semanticsegmentation_b200.dist import shard_range


def test_mean_matches_reference_semantics():
    assert rtda.mean([]) == 0                                   # lovasz_loss.py:33-37
    with pytest.raises(ValueError):
        rtda.mean([], empty="raise")
    assert rtda.mean([3.0]) == 3.0                              # single value returned undivided
    assert rtda.mean(iter([1.0, 2.0, 6.0])) == 3.0
    assert rtda.mean([1.0, float("nan"), 3.0], ignore_nan=True) == 2.0
    assert bool(rtda.isnan(torch.tensor(float("nan")))) and not bool(rtda.isnan(torch.tensor(1.0)))


def test_flatten_probas_layout():
    p = torch.arange(2 * 3 * 2 * 2, dtype=torch.float32).view(2, 3, 2, 2)
    l = torch.tensor([[[0, 1], [255, 2]], [[1, 1], [0, 255]]])
    flat, lab = rtda.flatten_probas(p, l)
    assert flat.shape == (8, 3) and lab.shape == (8,)
    assert torch.equal(flat[1], p[0, :, 0, 1])
    flat2, lab2 = rtda.flatten_probas(p, l, ignore=255)
    assert flat2.shape == (6, 3) and 255 not in lab2.tolist()
    crit = rtda.LovaszSoftmax(ignore=255)
    f3, l3 = crit.prob_flatten(p, l)                            # lovasz_loss.py:93-120
    assert torch.equal(f3, flat) and torch.equal(l3, lab)


def test_error_behaviour_matches_reference():
    crit = rtda.LovaszSoftmax(ignore=255)
    with pytest.raises(ValueError, match="Expected 4D input, but got 3D"):    # lovasz_loss.py:106-107
        crit(torch.rand(3, 4, 4), torch.zeros(4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU path"):
        crit(torch.rand(1, 3, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(ValueError):
        rtda.lovasz_softmax(torch.rand(1, 3, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long), classes="some")
    assert crit.reduction == "mean" and crit.ignore == 255 and list(crit.parameters()) == []


def test_per_class_iou_host(hist_golden):
    z = hist_golden
    for name in [str(n) for n in z["names"]]:
        got = rtda.per_class_iou(z[f"{name}/hist"].astype(np.float64))
        assert np.array_equal(got, z[f"{name}/iou"]), name
    assert rtda.mean_iou(z["val_slice/hist"].astype(np.float64)) == float(z["val_slice/miou"])
    # an absent class counts as IoU 0 in the mean (utils.py:112-114, validation.py:136)
    h = np.zeros((3, 3)); h[0, 0] = 5; h[1, 1] = 5
    assert rtda.mean_iou(h) == pytest.approx(2 / 3, rel=1e-5)  # the 1e-5 epsilon of utils.py:106


def test_synthetic_generators():
    p, l = synthetic.confident_case(2, 64, 96, seed=3)
    assert p.shape == (2, 19, 64, 96) and l.shape == (2, 64, 96) and l.dtype == torch.int64
    assert torch.allclose(p.sum(1), torch.ones(2, 64, 96), atol=1e-5)
    frac_ign = (l == 255).float().mean().item()
    assert 0.02 < frac_ign < 0.3
    valid = l != 255
    agree = (p.argmax(1)[valid] == l[valid]).float().mean().item()
    assert agree > 0.4                                           # trained-like
    lab, prd = synthetic.validation_maps(1, 64, 64, seed=4)
    v = lab != 255
    assert 0.7 < (lab[v] == prd[v]).float().mean().item() < 0.9
    p2, l2 = synthetic.confident_case(2, 64, 96, seed=3)
    assert torch.equal(p, p2) and torch.equal(l, l2)            # seeded


def test_shard_range_partitions():
    for n in (0, 1, 7, 16, 500):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_step_metrics_running_sums_and_periodic_read():
    """StepMetrics: the reference's running sums / periodic logging values
    (train_lovasz.py:133-158) without a per-step .item()."""
    import torch
    from rtda_semanticsegmentation_b200 import StepMetrics

    m = StepMetrics(("loss_total", "loss_ce", "loss_lovasz"), "cpu", every=3)
    want = {"loss_total": 0.0, "loss_ce": 0.0, "loss_lovasz": 0.0}
    reads = []
    for step in range(1, 8):
        ce, lv = torch.tensor(0.1 * step), torch.tensor(0.01 * step, requires_grad=True)
        tot = ce + 0.5 * lv
        m.update(loss_total=tot, loss_ce=ce, loss_lovasz=lv)
        for k, v in (("loss_total", tot), ("loss_ce", ce), ("loss_lovasz", lv)):
            want[k] += float(v)
        r = m.maybe_read(step)
        if r is not None:
            reads.append(step)
            assert r["steps"] == step and abs(r["loss_ce"] - 0.1 * step) < 1e-7
            assert abs(r["loss_total_sum"] - want["loss_total"]) < 1e-6
    assert reads == [3, 6]
    means = m.means()
    assert abs(means["loss_lovasz"] - want["loss_lovasz"] / 7) < 1e-7
    m.reset()
    assert m.read()["loss_ce_sum"] == 0.0 and m.steps == 0
