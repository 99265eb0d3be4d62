"""GPU development aid: a few confusion-matrix calls (1 map and 8 maps per call) for an ncu launch list.
usage: ncu --metrics gpu__time_duration.sum ... python scripts/gpu_cm_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtda_semanticsegmentation_b200 as rtda  # noqa: E402
from rtda_semanticsegmentation_b200 import synthetic  # noqa: E402

cm = rtda.ConfusionMatrix(19, 255, device="cuda")
for batch in (1, 8):
    maps = [synthetic.validation_maps(batch, 1024, 2048, seed=1234 + i, device="cuda") for i in range(12 // batch + 1)]
    torch.cuda.synchronize()
    for rep in range(3):
        for l, p in maps:
            cm.update(l, p)
    torch.cuda.synchronize()
lab8, prd8 = synthetic.validation_maps(1, 1024, 2048, seed=99, device="cuda")
cm.update(lab8.to(torch.uint8), prd8)          # uint8 labels, int64 predictions (9 B/px)
cm.update_from_logits(torch.randn(1, 19, 1024, 2048, device="cuda"), lab8)
torch.cuda.synchronize()
print("done", int(cm.compute().sum()))
