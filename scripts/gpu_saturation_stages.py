"""GPU development aid: per-kernel times (library events, eager) of the c2 shape as predictions saturate."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtda_semanticsegmentation_b200 as rtda  # noqa: E402
from rtda_semanticsegmentation_b200 import _lib, synthetic  # noqa: E402

for boost, scale in [(6.0, 3.0), (9.0, 3.0), (12.0, 3.0), (16.0, 3.0)]:
    lb = synthetic.blocky_labels(4, 512, 1024, seed=5, device="cuda")
    lg = synthetic.confident_logits(lb, seed=5, scale=scale, boost=boost)
    pr = torch.softmax(lg, 1)
    p = pr.clone().requires_grad_(True)
    for _ in range(3):
        rtda.lovasz_softmax(p, lb, ignore=255).backward()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    _lib.profile_read()
    for _ in range(10):
        rtda.lovasz_softmax(p, lb, ignore=255).backward()
    torch.cuda.synchronize()
    st = _lib.profile_read()
    _lib.profile_enable(False)
    g = p.grad
    zfrac = float((g == 0).float().mean())
    print(f"boost {boost}: " + " ".join(f"{k} {v[0] / max(v[1], 1) * 1e3:.1f}" for k, v in st.items() if v[0] / max(v[1], 1) * 1e3 > 5.0) + f"  zero gradients {zfrac:.3f}", flush=True)
