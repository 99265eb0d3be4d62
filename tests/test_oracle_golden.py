"""CPU: the oracle (torch restatement, integer restatement, C restatement) against
the golden vectors written from the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import golden_case
from oracle import c_oracle, hist_oracle, lovasz_oracle


def _names(z):
    return [str(n) for n in z["names"]]


def test_golden_files_present(lovasz_golden, hist_golden):
    assert len(_names(lovasz_golden)) >= 15
    assert len(_names(hist_golden)) >= 5


def test_lovasz_grad_known_answers(lovasz_golden):
    z = lovasz_golden
    for i in range(int(z["lovasz_grad/count"])):
        gt = z[f"lovasz_grad/{i}/gt"]
        want = z[f"lovasz_grad/{i}/out"]
        got = lovasz_oracle.jaccard_grad(torch.from_numpy(gt)).numpy()
        got_int = lovasz_oracle.jaccard_grad_int(gt)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert np.array_equal(got_int.view(np.uint32), want.view(np.uint32))
    # KAT-grad-1 of SURVEY.md section 4
    np.testing.assert_allclose(lovasz_oracle.jaccard_grad_int(np.array([1, 0, 1, 0, 0])),
                               [0.5, 0.166666627, 0.333333373, 0, 0], rtol=0, atol=1e-9)


@pytest.mark.parametrize("stable", [False, True])
def test_torch_oracle_matches_reference_fixtures(lovasz_golden, stable):
    torch.set_num_threads(1)
    for name in _names(lovasz_golden):
        c = golden_case(lovasz_golden, name)
        loss, grad = lovasz_oracle.lovasz_softmax_with_grad(
            torch.from_numpy(c["probas"]), torch.from_numpy(c["labels"]), c["ignore"], c["per_image"], stable=stable)
        assert abs(float(loss) - c["loss"]) <= 1e-6 * max(abs(c["loss"]), 1e-30), name
        if not c["ties"]:
            # no ties: stable and unstable orders coincide, gradients are reproducible bit for bit
            assert np.array_equal(grad.numpy(), c["grad"]), name
        else:
            # the reference's order inside tie groups is implementation-defined: only bound it
            assert np.abs(grad.numpy()).sum() == pytest.approx(np.abs(c["grad"]).sum(), rel=1e-4), name


def test_c_oracle_matches_fixtures_and_stable_oracle(lovasz_golden):
    for name in _names(lovasz_golden):
        c = golden_case(lovasz_golden, name)
        r = c_oracle.lovasz(c["probas"], c["labels"], c["ignore"], c["per_image"])
        assert abs(float(r["loss"]) - c["loss"]) <= 1e-6 * max(abs(c["loss"]), 1e-30), name
        _, g_stable = lovasz_oracle.lovasz_softmax_with_grad(
            torch.from_numpy(c["probas"]), torch.from_numpy(c["labels"]), c["ignore"], c["per_image"], stable=True)
        assert np.array_equal(r["grad"], g_stable.numpy()), name  # bit for bit, ties included
        if not c["ties"]:
            assert np.array_equal(r["grad"], c["grad"]), name


def test_c_oracle_grad_out_scaling(lovasz_golden):
    c = golden_case(lovasz_golden, "kat2_global")
    r1 = c_oracle.lovasz(c["probas"], c["labels"], c["ignore"], False, grad_out=1.0)
    r2 = c_oracle.lovasz(c["probas"], c["labels"], c["ignore"], False, grad_out=32768.0)
    assert np.array_equal(r2["grad"], r1["grad"] * np.float32(32768.0))  # exactly linear


def test_invariants_on_oracle(lovasz_golden):
    c = golden_case(lovasz_golden, "kat2_global")
    r = c_oracle.lovasz(c["probas"], c["labels"], c["ignore"], False)
    g = r["grad"]
    ign = c["labels"] == 255
    assert np.all(g.transpose(0, 2, 3, 1)[ign] == 0)                    # (i) zero at ignored pixels
    assert float(np.abs(g.astype(np.float64)).sum()) == pytest.approx(1.0, abs=1e-5)  # (iii) sum |grad| = 1
    onehot = np.zeros_like(g, dtype=bool)
    lab = np.where(ign, 0, c["labels"])
    np.put_along_axis(onehot, lab[:, None], True, axis=1)
    onehot &= ~ign[:, None]
    assert np.all(g[onehot] <= 0) and np.all(g[~onehot] >= 0)           # (ii) signs


def test_hist_oracle_matches_fixtures(hist_golden):
    z = hist_golden
    for name in _names(z):
        t, p, n = z[f"{name}/true"], z[f"{name}/pred"], int(z[f"{name}/n_class"])
        want = z[f"{name}/hist"]
        assert np.array_equal(hist_oracle.confusion_counts(t, p, n), want), name
        assert np.array_equal(c_oracle.confusion(t, p, n, ignore=None), want), name
        assert np.array_equal(hist_oracle.iou_per_class(want.astype(np.float64)), z[f"{name}/iou"]), name
    lab, prd = z["val_slice/labels"], z["val_slice/preds"]
    assert np.array_equal(hist_oracle.validation_confusion(lab, prd, 19, 255), z["val_slice/hist"])
    assert np.array_equal(c_oracle.confusion(lab, prd, 19, 255), z["val_slice/hist"])
    assert hist_oracle.mean_iou(z["val_slice/hist"].astype(np.float64)) == float(z["val_slice/miou"])


def test_c_oracle_against_midsize_reference_fixture(midsize_golden):
    """Mid-size cases (several buckets per segment on the GPU, saturated predictions with exact ties) recorded
    from the unmodified reference: loss, per-class sum of |gradient|, 4096 sampled gradient entries."""
    from conftest import assert_midsize, midsize_case
    from oracle import c_oracle

    for name in [str(n) for n in midsize_golden["names"]]:
        c = midsize_case(midsize_golden, name)
        r = c_oracle.lovasz(c["probas"].numpy(), c["labels"].numpy(), 255, c["per_image"])
        assert_midsize(c, float(r["loss"]), r["grad"], name)
