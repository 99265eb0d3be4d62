"""GPU development aid: final-gradient mode vs exact_backward on a full-size case, and where they differ."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtda_semanticsegmentation_b200 as rtda  # noqa: E402
from rtda_semanticsegmentation_b200 import synthetic  # noqa: E402

b, h, w = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (4, 512, 1024)))
per_image = len(sys.argv) > 4 and sys.argv[4] == "pi"
labels = synthetic.blocky_labels(b, h, w, seed=51, device="cuda")
logits = synthetic.confident_logits(labels, seed=51)
p = torch.softmax(logits, 1)
res = {}
for mode in (False, True):
    x = p.clone().requires_grad_(True)
    loss = rtda.LovaszSoftmax(ignore=255, per_image=per_image, exact_backward=mode)(x, labels)
    loss.backward()
    res[mode] = (float(loss), x.grad.clone())
d = (res[False][1] - res[True][1])
print("loss", res[False][0], res[True][0], "max |final - exact|", float(d.abs().max()), "of max", float(res[True][1].abs().max()),
      "differing elements", int((d != 0).sum()))
if (d != 0).any():
    idx = (d != 0).nonzero()[:10]
    for i in idx:
        bb, cc, yy, xx = [int(v) for v in i]
        print((bb, cc, yy, xx), "label", int(labels[bb, yy, xx]), "p", float(p[bb, cc, yy, xx]), "final", float(res[False][1][bb, cc, yy, xx]),
              "exact", float(res[True][1][bb, cc, yy, xx]))
    per_class = (d != 0).sum(dim=(0, 2, 3))
    print("per class", per_class.tolist(), "per image", (d != 0).sum(dim=(1, 2, 3)).tolist())
