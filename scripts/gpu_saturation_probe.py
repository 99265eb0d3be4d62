"""How the bucket path behaves as predictions saturate: D_conf with growing boost at the c2 shape.
Prints fallback segments and fwd+bwd time."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtda_semanticsegmentation_b200 as rtda  # noqa: E402
from rtda_semanticsegmentation_b200 import synthetic  # noqa: E402

for boost, scale in [(6.0, 3.0), (9.0, 3.0), (12.0, 3.0), (16.0, 3.0), (12.0, 1.5), (20.0, 2.0)]:
    lb = synthetic.blocky_labels(4, 512, 1024, seed=5, device="cuda")
    lg = synthetic.confident_logits(lb, seed=5, scale=scale, boost=boost)
    pr = torch.softmax(lg, 1)
    p = pr.clone().requires_grad_(True)
    for _ in range(3):
        rtda.lovasz_softmax(p, lb, ignore=255).backward()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        rtda.lovasz_softmax(p, lb, ignore=255).backward()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    fb = int(rtda.segment_sort_paths((4, 19, 512, 1024), False).sum())
    zero = float((pr == 1.0).float().mean() * 19)
    print(f"boost {boost} scale {scale}: {dt * 1e3:.3f} ms/step, lsd segments {fb}/19, pixels with p_max == 1.0: {zero:.3f}", flush=True)
