"""Debug aid: one forward of a given shape, then the bucket-path bookkeeping (flags, buckets, occupancy)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtda_semanticsegmentation_b200 as rtda  # noqa: E402
from rtda_semanticsegmentation_b200 import _lib, synthetic  # noqa: E402
from rtda_semanticsegmentation_b200.lovasz_loss import _workspaces  # noqa: E402

b, c, h, w = [int(x) for x in sys.argv[1:5]]
per_image = len(sys.argv) > 5 and sys.argv[5] == "pi"
pr, lb = synthetic.confident_case(b, h, w, n_class=c, seed=51, device="cuda")
loss = rtda.lovasz_softmax(pr.requires_grad_(True), lb, ignore=255, per_image=per_image)
loss.backward()
torch.cuda.synchronize()
o = (ctypes.c_size_t * 24)()
_lib.load().rtda_lovasz_debug_layout(b, c, h * w, int(per_image), o)
ws = next(iter(_workspaces.values()))
S, NBK = int(o[14]), int(o[22])
fb = ws[int(o[16]):int(o[16]) + 4 * S].view(torch.int32).cpu().numpy()
nbs = ws[int(o[17]):int(o[17]) + 4 * S].view(torch.int32).cpu().numpy() & 0x7fffffff
cur = ws[int(o[18]):int(o[18]) + 8 * S * NBK].view(torch.int64).view(S, NBK).cpu().numpy()
occ = cur & 0xffffffff
print("loss", float(loss), "S", S, "NBK", NBK, "fallback flags", fb.tolist()[:40], "buckets", nbs.tolist()[:40])
for s in range(min(S, 6)):
    o_ = occ[s, :max(nbs[s], 1)]
    top = np.argsort(-o_)[:5]
    print(f"seg {s}: nb {nbs[s]} sum {o_.sum()} max {o_.max()} mean {o_.mean():.0f} largest {[(int(i), int(o_[i])) for i in top]} first {o_[:3].tolist()} last {o_[-3:].tolist()}")
