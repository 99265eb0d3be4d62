"""The drop-in seam: the reference's import paths resolve to the B200 implementation."""
import os
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

import rtda_semanticsegmentation_b200 as rtda
from rtda_semanticsegmentation_b200 import dropin, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_path_shadowing_resolves_reference_import_lines():
    code = ("import sys; sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2]);"
            "from losses.lovasz_loss import LovaszSoftmax, lovasz_grad, mean, isnan;"   # main.py:27
            "import rtda_semanticsegmentation_b200 as r;"
            "assert LovaszSoftmax is r.LovaszSoftmax and mean is r.mean; print('ok')")
    out = subprocess.run([sys.executable, "-c", code, ROOT, os.path.join(ROOT, "rtda_semanticsegmentation_b200", "dropin")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_install_rebinds_modules(monkeypatch):
    fake_utils = types.ModuleType("utils")
    fake_utils.fast_hist = lambda *a: None
    fake_utils.per_class_iou = lambda *a: None
    fake_val = types.ModuleType("validation")
    fake_val.fast_hist = fake_utils.fast_hist
    fake_val.per_class_iou = fake_utils.per_class_iou
    fake_losses = types.ModuleType("losses")
    fake_ll = types.ModuleType("losses.lovasz_loss")
    fake_ll.LovaszSoftmax = object
    monkeypatch.setitem(sys.modules, "utils", fake_utils)
    monkeypatch.setitem(sys.modules, "validation", fake_val)
    monkeypatch.setitem(sys.modules, "losses", fake_losses)
    monkeypatch.setitem(sys.modules, "losses.lovasz_loss", fake_ll)
    done = dropin.install()
    assert fake_ll.LovaszSoftmax is rtda.LovaszSoftmax
    assert fake_val.fast_hist is rtda.fast_hist and fake_utils.per_class_iou is rtda.per_class_iou
    assert "losses.lovasz_loss.LovaszSoftmax" in done and "validation.fast_hist" in done


@pytest.mark.gpu
def test_training_step_call_pattern_matches_reference_step():
    """train_lovasz.py:98-112 (AMP branch) with the drop-in criterion, against the same step
    computed with the oracle restatement of the reference loss on the CPU."""
    from oracle import lovasz_oracle

    _, labels = synthetic.confident_case(2, 32, 48, seed=31)
    torch.manual_seed(0)
    logits0 = torch.randn(2, 19, 32, 48)
    # reference-side step on CPU (fp32): CE + 0.5 * Lovasz(softmax(logits)), backward to the logits
    lg = logits0.clone().requires_grad_(True)
    ce = torch.nn.CrossEntropyLoss(ignore_index=255)(lg, labels)
    lv = lovasz_oracle.lovasz_softmax(torch.softmax(lg, 1), labels, ignore=255, stable=True)
    (ce + 0.5 * lv).backward()
    # drop-in step on the GPU, call for call as train_lovasz.py:83-112
    criterion_ce = torch.nn.CrossEntropyLoss(ignore_index=255)
    criterion_lovasz = rtda.LovaszSoftmax(ignore=255)                     # main.py:299-302
    scaler = torch.amp.GradScaler("cuda")
    gl = logits0.cuda().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        loss_ce = criterion_ce(gl, labels.cuda())                          # :98
        probas = torch.nn.functional.softmax(gl, dim=1)                    # :101
        loss_lovasz = criterion_lovasz(probas, labels.cuda())              # :102
        loss_total = loss_ce + 0.5 * loss_lovasz                           # :105
    scaler.scale(loss_total).backward()                                    # :108
    torch.cuda.synchronize()
    assert loss_lovasz.item() == pytest.approx(float(lv), rel=2e-6)       # :135 .item() works
    got = gl.grad.cpu() / scaler.get_scale()
    assert torch.allclose(got, lg.grad, rtol=1e-4, atol=1e-7)


@pytest.mark.gpu
def test_validation_slice_call_pattern():
    """validation.py:95-109,135-136 with the drop-in fast_hist / per_class_iou and with the
    device-resident accumulator: identical confusion matrix and mIoU."""
    from oracle import hist_oracle

    labels, _ = synthetic.validation_maps(1, 64, 80, seed=33)
    outputs = synthetic.confident_logits(labels, seed=33).cuda()
    preds = torch.argmax(outputs, dim=1)                                   # :95
    labels_np, preds_np = labels.numpy(), preds.cpu().numpy()              # :98-99
    valid = labels_np != 255                                               # :102
    hist = np.zeros((19, 19))
    hist += rtda.fast_hist(labels_np[valid].flatten(), preds_np[valid].flatten(), 19)   # :105-109
    want = hist_oracle.validation_confusion(labels_np, preds_np, 19, 255)
    assert np.array_equal(hist, want)
    cm = rtda.ConfusionMatrix(19, 255).update_from_logits(outputs, labels.cuda())
    assert np.array_equal(cm.compute(), want)
    assert float(np.nanmean(rtda.per_class_iou(hist))) == hist_oracle.mean_iou(want.astype(np.float64))   # :135-136
