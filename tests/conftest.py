"""pytest configuration: `gpu` marker, repo root on sys.path, shared helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly on a box without CUDA only when explicitly selected;
    in the default CPU run (-m "not gpu") they are deselected by the marker."""
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    markexpr = config.getoption("-m") or ""
    if "gpu" in markexpr and "not gpu" not in markexpr:
        return  # the user asked for GPU tests: let them fail rather than hide them
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lovasz_golden():
    z = np.load(os.path.join(GOLDEN, "lovasz_golden.npz"))
    return z


@pytest.fixture(scope="session")
def hist_golden():
    return np.load(os.path.join(GOLDEN, "hist_golden.npz"))


def golden_case(z, name):
    ig = int(z[f"{name}/ignore"])
    return dict(
        probas=z[f"{name}/probas"], labels=z[f"{name}/labels"], loss=float(z[f"{name}/loss"]),
        grad=z[f"{name}/grad"], ignore=None if ig < 0 else ig, per_image=bool(z[f"{name}/per_image"]),
        ties=bool(z[f"{name}/ties"]), ref_tensor=bool(z[f"{name}/ref_returned_tensor"]))


@pytest.fixture(scope="session")
def midsize_golden():
    return np.load(os.path.join(GOLDEN, "lovasz_midsize_golden.npz"))


def midsize_case(z, name):
    """Regenerates the inputs of a mid-size reference case (tests/golden/make_golden.py midsize_cases) and
    returns them with what the unmodified reference recorded.  The inputs are not stored; their checksum is."""
    import torch

    from rtda_semanticsegmentation_b200 import synthetic

    b, h, w, seed, per_image = [int(v) for v in z[f"{name}/spec"]]
    scale, boost = [float(v) for v in z[f"{name}/scale_boost"]]
    labels = synthetic.blocky_labels(b, h, w, seed=seed)
    probas = torch.softmax(synthetic.confident_logits(labels, seed=seed, scale=scale, boost=boost), 1)
    chk = [int(probas.view(torch.int32).to(torch.int64).sum()), int(labels.sum())]
    if chk != [int(v) for v in z[f"{name}/input_checksum"]]:
        pytest.skip(f"{name}: this torch build generates different synthetic inputs than the fixture was written from")
    return dict(probas=probas, labels=labels, per_image=bool(per_image), loss=float(z[f"{name}/loss"]),
                abs_grad_per_class=z[f"{name}/abs_grad_per_class"], grad_idx=z[f"{name}/grad_idx"],
                grad_val=z[f"{name}/grad_val"], grad_max=float(z[f"{name}/grad_max"]),
                ties=int(z[f"{name}/exact_one_probabilities"]) > 0)


def assert_midsize(case, loss, grad, name):
    """loss within 1e-5 of the reference; the per-class sums of |gradient| (invariant under the order inside
    a group of equal errors) to 1e-9; the sampled gradient entries within 1e-5 of the largest gradient -
    bit for bit where no exact ties exist.  Inside tie groups (saturated predictions) the reference's own
    order is whatever its unstable sort produced: there only a small fraction may differ, by little."""
    assert abs(loss - case["loss"]) <= 1e-5 * abs(case["loss"]), name
    apc = np.abs(grad.astype(np.float64)).sum(axis=(0, 2, 3))
    assert np.abs(apc - case["abs_grad_per_class"]).max() <= 1e-9 * max(case["abs_grad_per_class"].max(), 1e-30), name
    got = grad.ravel()[case["grad_idx"]]
    d = np.abs(got - case["grad_val"])
    if not case["ties"]:
        assert d.max() <= 1e-5 * case["grad_max"], name
        assert np.array_equal(got, case["grad_val"]), name
    else:
        assert (d > 1e-5 * case["grad_max"]).mean() <= 0.02 and d.max() <= 2e-4 * case["grad_max"], (name, d.max() / case["grad_max"])
