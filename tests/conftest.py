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
