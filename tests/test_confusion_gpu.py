"""GPU: the confusion-matrix kernels (through the C ABI) are bit-exact against the oracle and
the golden vectors written from the reference's fast_hist."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, hist_oracle
import rtda_semanticsegmentation_b200 as rtda
from rtda_semanticsegmentation_b200 import synthetic

pytestmark = pytest.mark.gpu


def test_golden_fast_hist(hist_golden):
    z = hist_golden
    for name in [str(n) for n in z["names"]]:
        t, p, n = z[f"{name}/true"], z[f"{name}/pred"], int(z[f"{name}/n_class"])
        got = rtda.fast_hist(t, p, n)                       # reference signature: numpy in, numpy out
        assert got.dtype == np.int64 and got.shape == (n, n)
        assert np.array_equal(got, z[f"{name}/hist"]), name
        assert np.array_equal(rtda.per_class_iou(got.astype(np.float64)), z[f"{name}/iou"]), name
    lab, prd = torch.from_numpy(z["val_slice/labels"]).cuda(), torch.from_numpy(z["val_slice/preds"]).cuda()
    h = rtda.confusion_matrix(lab, prd, 19, 255)
    assert h.is_cuda and h.dtype == torch.int64
    assert np.array_equal(h.cpu().numpy(), z["val_slice/hist"])
    assert rtda.mean_iou(h) == float(z["val_slice/miou"])


@pytest.mark.parametrize("n_class", [2, 19, 40, 64, 150, 300])
@pytest.mark.parametrize("n", [1, 31, 1000, 100003])
def test_random_vs_oracle(n_class, n):
    r = np.random.default_rng(n_class * 7 + n)
    t = r.integers(-1, n_class + 1, size=n)
    p = r.integers(-1, n_class + 1, size=n)
    t[r.random(n) < 0.1] = 255
    got = rtda.confusion_matrix(torch.from_numpy(t).cuda(), torch.from_numpy(p).cuda(), n_class, 255).cpu().numpy()
    assert np.array_equal(got, hist_oracle.validation_confusion(t, p, n_class, 255))
    assert np.array_equal(got, c_oracle.confusion(t, p, n_class, 255))


@pytest.mark.parametrize("ldt,pdt", [(torch.uint8, torch.int64), (torch.int32, torch.int32), (torch.uint8, torch.uint8),
                                     (torch.int64, torch.uint8)])
def test_dtype_variants(ldt, pdt):
    lab, prd = synthetic.validation_maps(2, 100, 130, seed=3)
    want = hist_oracle.validation_confusion(lab.numpy(), prd.numpy(), 19, 255)
    got = rtda.confusion_matrix(lab.to(ldt).cuda(), prd.to(pdt).cuda(), 19, 255)
    assert np.array_equal(got.cpu().numpy(), want)


def test_accumulator_and_empty_input():
    cm = rtda.ConfusionMatrix(19, 255)
    total = np.zeros((19, 19), np.int64)
    for i in range(3):
        lab, prd = synthetic.validation_maps(1, 64, 96, seed=10 + i)
        cm.update(lab.cuda(), prd.cuda())
        total += hist_oracle.validation_confusion(lab.numpy(), prd.numpy(), 19, 255)
    cm.update(torch.zeros(0, dtype=torch.long, device="cuda"), torch.zeros(0, dtype=torch.long, device="cuda"))
    assert np.array_equal(cm.compute(), total)
    assert cm.mean_iou() == hist_oracle.mean_iou(total.astype(np.float64))
    assert np.array_equal(cm.per_class_iou(), hist_oracle.iou_per_class(total.astype(np.float64)))
    cm.reset()
    assert not cm.compute().any()


def test_argmax_fused_matches_torch_argmax():
    lab, _ = synthetic.validation_maps(2, 70, 90, seed=5)
    logits = synthetic.confident_logits(lab, seed=5).cuda()
    logits[0, 3, 0, :5] = logits[0, 7, 0, :5] = 50.0        # exact ties: the first maximum wins, like torch.argmax
    preds = torch.argmax(logits, dim=1)                     # validation.py:95
    want = hist_oracle.validation_confusion(lab.numpy(), preds.cpu().numpy(), 19, 255)
    got = rtda.confusion_matrix_from_logits(logits, lab.cuda(), 255)
    assert np.array_equal(got.cpu().numpy(), want)
    got8 = rtda.confusion_matrix_from_logits(logits, lab.to(torch.uint8).cuda(), 255)
    assert np.array_equal(got8.cpu().numpy(), want)


def test_full_size_validation_map_and_checksum():
    """BASELINE.json configs[3] shape: 1024x2048 maps.  Bit-exact vs the C oracle; the grand
    total equals the number of valid pixels (a size-independent checksum)."""
    lab, prd = synthetic.validation_maps(2, 1024, 2048, seed=1234 + 400, device="cuda")
    h = rtda.confusion_matrix(lab, prd, 19, 255).cpu().numpy()
    assert np.array_equal(h, c_oracle.confusion(lab.cpu().numpy(), prd.cpu().numpy(), 19, 255))
    assert h.sum() == int((lab != 255).sum())
    assert np.array_equal(h.sum(1), np.bincount(lab[lab != 255].cpu().numpy(), minlength=19))


@pytest.mark.parametrize("tail", [0, 1, 15, 16, 2049])
def test_bulk_copy_staged_kernel_on_a_batched_call(tail):
    """Calls of 8 chunks per SM or more take k_confusion_tma (maps staged in shared memory by bulk copies).  Three
    1024x2048 maps plus a ragged tail: the last chunk is partial, its end not a multiple of the 16 pixels a copy moves."""
    lab, prd = synthetic.validation_maps(3, 1024, 2048, seed=77, device="cuda")
    n = 3 * 1024 * 2048 - 4096 + tail
    lab, prd = lab.reshape(-1)[:n].contiguous(), prd.reshape(-1)[:n].contiguous()
    prd[5] = 19; prd[6] = -1; lab[7] = 300; lab[8] = -7            # out-of-range values count nothing (utils.py:78-83)
    want = c_oracle.confusion(lab.cpu().numpy(), prd.cpu().numpy(), 19, 255)
    assert np.array_equal(rtda.confusion_matrix(lab, prd, 19, 255).cpu().numpy(), want)
    # narrower storage types of either map (9 and 5 bytes per pixel)
    lab8 = torch.where((lab >= 0) & (lab < 256), lab, torch.full_like(lab, 255)).to(torch.uint8)
    want8 = c_oracle.confusion(lab8.cpu().numpy().astype(np.int64), prd.cpu().numpy(), 19, 255)
    assert np.array_equal(rtda.confusion_matrix(lab8, prd, 19, 255).cpu().numpy(), want8)
    assert np.array_equal(rtda.confusion_matrix(lab8, prd.to(torch.int32), 19, 255).cpu().numpy(), want8)
