"""GPU: the fused CE + Lovasz path from logits against the reference-side
composition  CE(logits) + w * Lovasz(softmax(logits))  computed on the CPU with torch and the oracle.

Tolerance: the fused softmax differs from torch's by rounding (1 ulp on a probability can swap two
neighbouring ranks), so this path is held to 1e-5 relative on the losses and 1e-5 of the largest
gradient magnitude elementwise - the north_star's floating-point bar - not to bit equality."""
import numpy as np
import pytest
import torch

from oracle import lovasz_oracle
import rtda_semanticsegmentation_b200 as rtda
from rtda_semanticsegmentation_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["bucket", "lsd"])
def sort_path(request):
    rtda.set_sort_path(request.param)
    yield request.param
    rtda.set_sort_path("bucket")


def reference_step(logits, labels, weight, per_image, go=1.0):
    lg = logits.clone().requires_grad_(True)
    ce = torch.nn.CrossEntropyLoss(ignore_index=255)(lg, labels)                 # train_lovasz.py:98
    lv = lovasz_oracle.lovasz_softmax(torch.softmax(lg, 1), labels, ignore=255, per_image=per_image, stable=True)  # :101-102
    if not torch.is_tensor(lv):
        lv = torch.zeros(())
    total = ce + weight * lv                                                     # :105
    (total * go).backward()
    return float(total), float(ce), float(lv), lg.grad


@pytest.mark.parametrize("shape,per_image", [((2, 19, 32, 48), False), ((3, 19, 33, 47), True), ((1, 5, 7, 13), False),
                                             ((2, 19, 64, 128), False), ((4, 7, 16, 20), True)])
def test_hybrid_matches_reference_composition(shape, per_image):
    b, c, h, w = shape
    g = torch.Generator().manual_seed(sum(shape))
    logits = torch.randn(shape, generator=g) * 3
    labels = torch.randint(0, c, (b, h, w), generator=g)
    labels[torch.rand((b, h, w), generator=g) < 0.15] = 255
    want_total, want_ce, want_lv, want_grad = reference_step(logits, labels, 0.5, per_image, go=2.0)
    gl = logits.cuda().requires_grad_(True)
    total, ce, lv = rtda.hybrid_ce_lovasz(gl, labels.cuda(), ignore=255, lovasz_weight=0.5, per_image=per_image)
    (total * 2.0).backward()
    torch.cuda.synchronize()
    assert float(ce) == pytest.approx(want_ce, rel=1e-5)
    assert float(lv) == pytest.approx(want_lv, rel=1e-5)
    assert float(total) == pytest.approx(want_total, rel=1e-5)
    got = gl.grad.cpu()
    assert float((got - want_grad).abs().max()) <= 1e-5 * float(want_grad.abs().max())
    assert not got.permute(0, 2, 3, 1)[labels == 255].any()                     # ignored pixels get no gradient


def test_hybrid_module_under_autocast_and_gradscaler():
    """The AMP branch of train_lovasz.py:83-112 with the fused criterion."""
    _, labels = synthetic.confident_case(2, 32, 48, seed=41)
    torch.manual_seed(1)
    logits = torch.randn(2, 19, 32, 48)
    want_total, want_ce, want_lv, want_grad = reference_step(logits, labels, 0.5, False)
    crit = rtda.HybridCELovasz(ignore=255, lovasz_weight=0.5)
    scaler = torch.amp.GradScaler("cuda")
    gl = logits.cuda().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        loss_total = crit(gl, labels.cuda())
    scaler.scale(loss_total).backward()
    torch.cuda.synchronize()
    assert loss_total.item() == pytest.approx(want_total, rel=1e-5)
    assert crit.last_ce.item() == pytest.approx(want_ce, rel=1e-5) and crit.last_lovasz.item() == pytest.approx(want_lv, rel=1e-5)
    got = gl.grad.cpu() / scaler.get_scale()
    assert float((got - want_grad).abs().max()) <= 2e-5 * float(want_grad.abs().max())


def test_hybrid_equals_unfused_b200_path_full_size():
    """c2 shape: fused path vs softmax -> LovaszSoftmax + CrossEntropyLoss on the same GPU."""
    labels = synthetic.blocky_labels(4, 512, 1024, seed=51, device="cuda")
    logits = synthetic.confident_logits(labels, seed=51)
    a = logits.clone().requires_grad_(True)
    total, ce, lv = rtda.hybrid_ce_lovasz(a, labels, 255, 0.5)
    total.backward()
    b = logits.clone().requires_grad_(True)
    ce2 = torch.nn.functional.cross_entropy(b, labels, ignore_index=255)
    lv2 = rtda.LovaszSoftmax(ignore=255)(torch.softmax(b, 1), labels)
    (ce2 + 0.5 * lv2).backward()
    torch.cuda.synchronize()
    assert float(ce) == pytest.approx(float(ce2), rel=1e-5) and float(lv) == pytest.approx(float(lv2), rel=1e-5)
    assert float((a.grad - b.grad).abs().max()) <= 1e-5 * float(b.grad.abs().max())
    assert not a.grad.permute(0, 2, 3, 1)[labels == 255].any()


def test_hybrid_label_dtypes_and_errors():
    logits = torch.randn(1, 4, 6, 6, device="cuda", requires_grad=True)
    labels = torch.randint(0, 4, (1, 6, 6), device="cuda")
    t64 = rtda.hybrid_ce_lovasz(logits, labels)[0]
    t8 = rtda.hybrid_ce_lovasz(logits, labels.to(torch.uint8))[0]
    assert float(t64) == float(t8)
    with pytest.raises(ValueError, match="Expected 4D input"):
        rtda.hybrid_ce_lovasz(logits[0], labels)
    with pytest.raises(RuntimeError, match="no CPU path"):
        rtda.hybrid_ce_lovasz(logits.cpu(), labels.cpu())
