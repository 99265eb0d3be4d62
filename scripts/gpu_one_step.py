"""Runs a few fwd+bwd steps of one workload (for ncu captures).  usage: gpu_one_step.py [c2|c3|c1] [steps] [path]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtda_semanticsegmentation_b200 as rtda  # noqa: E402
from rtda_semanticsegmentation_b200 import synthetic  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if len(sys.argv) > 3:
    rtda.set_sort_path(sys.argv[3])
shape = {"c1": (1, 512, 1024, False), "c2": (4, 512, 1024, False), "c3": (4, 720, 1280, True),
         "small": (2, 128, 256, False), "small_pi": (3, 96, 160, True), "ties": (2, 128, 256, False)}[wl]
pr, lb = synthetic.confident_case(shape[0], shape[1], shape[2], seed=11, device="cuda")
if wl == "ties":            # one class constant (massive ties -> LSD fallback for it), twins elsewhere (slot sharing)
    pr[:, 3] = 0.125
    pr[..., 1::2] = pr[..., 0::2]
    lb[..., 1::2] = lb[..., 0::2]
p = pr.requires_grad_(True)
for _ in range(steps):
    loss = rtda.lovasz_softmax(p, lb, ignore=255, per_image=shape[3])
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss), "lsd segments", int(rtda.segment_sort_paths((shape[0], 19, shape[1], shape[2]), shape[3]).sum()))
