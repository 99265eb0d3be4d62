"""GPU development aid: a few eager fwd+bwd steps of per-image c2 (final-gradient mode) for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtda_semanticsegmentation_b200 as rtda  # noqa: E402
from rtda_semanticsegmentation_b200 import synthetic  # noqa: E402

p, l = synthetic.confident_case(4, 512, 1024, seed=1234, device="cuda")
crit = rtda.LovaszSoftmax(ignore=255, per_image=True)
for i in range(4):
    x = p.detach().requires_grad_(True)
    crit(x, l).backward()
torch.cuda.synchronize()
print("ok")
