"""GPU development aid: eager (no CUDA graph, no profiling events) fwd+bwd step time of the public modules,
next to the host's enqueue time per step, for the plain and the fused hybrid loss.
usage (GPU box): python scripts/gpu_eager_probe.py [c2|c1|c3|c5] [steps]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import rtda_semanticsegmentation_b200 as rtda  # noqa: E402
from rtda_semanticsegmentation_b200 import synthetic  # noqa: E402

SHAPES = {"c1": (1, 512, 1024, False), "c2": (4, 512, 1024, False), "c3": (16, 720, 1280, True), "c5": (8, 512, 1024, False)}


def run(name, fn, sets, steps):
    for i in range(5):
        fn(*sets[i % len(sets)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for i in range(steps):
        fn(*sets[i % len(sets)])
    host = (time.perf_counter() - t0) * 1e3 / steps
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:28s} eager {e0.elapsed_time(e1) / steps:.4f} ms/step, host enqueue {host:.4f} ms/step", flush=True)


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    b, h, w, per_image = SHAPES[wl]
    sets, lsets = [], []
    for i in range(3 if wl != "c3" else 2):
        l = synthetic.blocky_labels(b, h, w, seed=1234 + 1000 * i, device="cuda")
        x = synthetic.confident_logits(l, seed=1234 + 1000 * i)
        lsets.append((x, l))
        sets.append((torch.softmax(x, 1), l))
    lov = rtda.LovaszSoftmax(ignore=255, per_image=per_image)
    hyb = rtda.HybridCELovasz(ignore=255, per_image=per_image)

    def plain(p, l):
        p = p.detach().requires_grad_(True)
        lov(p, l).backward()

    def fwd_only(p, l):
        with torch.no_grad():
            lov(p, l)

    def fused(x, l):
        x = x.detach().requires_grad_(True)
        hyb(x, l).backward()

    def unfused(x, l):
        x = x.detach().requires_grad_(True)
        (torch.nn.functional.cross_entropy(x, l, ignore_index=255) + 0.5 * lov(torch.softmax(x, 1), l)).backward()

    run(f"{wl} lovasz fwd+bwd", plain, sets, steps)
    run(f"{wl} lovasz fwd (no_grad)", fwd_only, sets, steps)
    run(f"{wl} hybrid fused fwd+bwd", fused, lsets, steps)
    run(f"{wl} hybrid unfused fwd+bwd", unfused, lsets, steps)


if __name__ == "__main__":
    main()
