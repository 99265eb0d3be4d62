"""GPU development aid: graph-replayed fwd+bwd time of ONE build of the library (RTDA_LIB_PATH) on a
BASELINE workload, its per-launch stage times, and a checksum of loss + gradient so that kernel
variants can be compared with the default build bit for bit.  usage (GPU box):
    RTDA_LIB_PATH=$PWD/variants/x.so python scripts/gpu_variant_time.py [c2|c1|c3|c5] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import rtda_semanticsegmentation_b200 as rtda  # noqa: E402
from rtda_semanticsegmentation_b200 import _lib, synthetic  # noqa: E402

SHAPES = {"c1": (1, 512, 1024, False), "c2": (4, 512, 1024, False), "c2pi": (4, 512, 1024, True), "c3": (16, 720, 1280, True), "c5": (8, 512, 1024, False)}


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    b, h, w, per_image = SHAPES[wl]
    name = os.path.basename(os.environ.get("RTDA_LIB_PATH", "default"))
    n_sets = 3 if wl != "c3" else 2
    sets = [synthetic.confident_case(b, h, w, seed=1234 + 1000 * i, device="cuda") for i in range(n_sets)]
    crit = rtda.LovaszSoftmax(ignore=255, per_image=per_image, exact_backward=bool(int(os.environ.get("RTDA_EXACT_BACKWARD", "0"))))
    graphs = []
    side = torch.cuda.Stream()
    for p, l in sets:
        pin = p.detach().requires_grad_(True)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                crit(pin, l).backward()
                pin.grad = None
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss = crit(pin, l)
            loss.backward()
        graphs.append((g, loss, pin))
    for i in range(6):
        graphs[i % n_sets][0].replay()
    torch.cuda.synchronize()
    best = 1e9
    tot = 0.0
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            graphs[i % n_sets][0].replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        best = min(best, ms)
        tot += ms
    g, loss, pin = graphs[0]
    g.replay()
    torch.cuda.synchronize()
    csum = int(pin.grad.view(torch.int32).to(torch.int64).sum().item())
    lbits = loss.detach().view(torch.int32).item()
    # per-launch stage times (eager launches on the default stream, library events)
    _lib.profile_enable(True)
    _lib.profile_read()
    for i in range(10):
        p = sets[i % n_sets][0].detach().requires_grad_(True)
        crit(p, sets[i % n_sets][1]).backward()
    torch.cuda.synchronize()
    st = _lib.profile_read()
    _lib.profile_enable(False)
    fb = int(rtda.segment_sort_paths((b, 19, h, w), per_image).sum())
    pick = {k: v[0] / max(v[1], 1) * 1e3 for k, v in st.items()}
    print(f"{wl} {name:12s} best {best:.4f} mean {tot / 3:.4f} ms/step  {b * h * w / best / 1e3:.0f} Mpix/s  fb {fb}  loss_bits {lbits:#x} grad_csum {csum}  "
          f"prep {pick.get('k_prep', 0):.1f} fin {pick.get('k_finalize', 0):.1f} split {pick.get('k_split', 0):.1f} plan {pick.get('k_plan', 0):.1f} part {pick.get('k_part', 0):.1f} local {pick.get('k_local', 0):.1f} bwd {pick.get('k_bwd', 0):.1f} us", flush=True)


if __name__ == "__main__":
    main()
