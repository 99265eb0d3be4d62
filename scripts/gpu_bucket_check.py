"""GPU development check of the bucket path: parity of both sort paths with the C oracle on a
range of shapes / distributions, which segments fell back, and per-stage times on the c2 shape.
Development aid (the build container has no GPU); the product tests are tests/test_*_gpu.py."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import c_oracle  # noqa: E402
import rtda_semanticsegmentation_b200 as rtda  # noqa: E402
from rtda_semanticsegmentation_b200 import _lib, synthetic  # noqa: E402


def run(probas, labels, per_image, path):
    rtda.set_sort_path(path)
    p = probas.cuda().requires_grad_(True)
    loss = rtda.lovasz_softmax(p, labels.cuda(), per_image=per_image, ignore=255)
    loss.backward()
    torch.cuda.synchronize()
    fb = rtda.segment_sort_paths(tuple(probas.shape), per_image)
    return float(loss), p.grad.cpu().numpy(), fb.cpu().numpy()


def check(name, probas, labels, per_image=False):
    want = c_oracle.lovasz(probas.numpy(), labels.numpy(), 255, per_image)
    ok = True
    for path in ("bucket", "lsd"):
        loss, grad, fb = run(probas, labels, per_image, path)
        lok = abs(loss - float(want["loss"])) <= 2e-6 * max(abs(float(want["loss"])), 1e-30)
        gok = np.array_equal(grad, want["grad"])
        nbad = int((grad != want["grad"]).sum())
        print(f"{name:34s} {path:6s} loss {loss:.7f} want {float(want['loss']):.7f} {'ok' if lok else 'LOSS-MISMATCH'} "
              f"grad {'bit-exact' if gok else f'MISMATCH({nbad})'} lsd-segments {int(fb.sum())}/{fb.size}", flush=True)
        ok &= lok and gok
    return ok


def main():
    ok = True
    g = torch.Generator().manual_seed(3)
    for shape in [(1, 19, 16, 24), (2, 19, 37, 53), (1, 2, 1, 1), (2, 19, 128, 256), (1, 19, 300, 500), (3, 7, 200, 333)]:
        b, c, h, w = shape
        pr = torch.softmax(torch.randn(shape, generator=g) * 2.5, 1)
        lb = torch.randint(0, c, (b, h, w), generator=g)
        lb[torch.rand((b, h, w), generator=g) < 0.12] = 255
        for per_image in (False, True):
            ok &= check(f"randn {shape} pi={int(per_image)}", pr, lb, per_image)
    pr, lb = synthetic.confident_case(2, 512, 512, seed=5)
    ok &= check("D_conf 2x19x512x512", pr, lb)
    ok &= check("D_conf 2x19x512x512 per-image", pr, lb, True)
    pr, lb = synthetic.diffuse_case(2, 512, 512, seed=6)
    ok &= check("D_diffuse 2x19x512x512", pr, lb)
    lb5 = torch.randint(0, 5, (2, 200, 200), generator=g)
    ok &= check("uniform 0.2 (all tied)", torch.full((2, 5, 200, 200), 0.2), lb5)
    onehot = torch.zeros(2, 5, 200, 200).scatter_(1, torch.randint(0, 5, (2, 1, 200, 200), generator=g), 1.0)
    ok &= check("one-hot (errors 0 and 1)", onehot, lb5)
    quant = torch.randint(0, 5, (2, 5, 200, 200), generator=g).float() / 4
    ok &= check("quantised", quant, lb5)
    sat = torch.softmax(torch.randn((1, 19, 256, 512), generator=g) * 30, 1)
    lbs = torch.randint(0, 19, (1, 256, 512), generator=g)
    ok &= check("saturated softmax (x30)", sat, lbs)
    print("ALL OK" if ok else "FAILURES", flush=True)

    # timing on the c2 shape, both paths
    pr, lb = synthetic.confident_case(4, 512, 1024, seed=11, device="cuda")
    for path in ("bucket", "lsd"):
        rtda.set_sort_path(path)
        p = pr.clone().requires_grad_(True)
        for _ in range(3):
            rtda.lovasz_softmax(p, lb, ignore=255).backward()
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        t0 = time.perf_counter()
        for _ in range(10):
            rtda.lovasz_softmax(p, lb, ignore=255).backward()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        prof = _lib.profile_read()
        _lib.profile_enable(False)
        fb = rtda.segment_sort_paths((4, 19, 512, 1024), False)
        if path == "bucket":
            import ctypes
            o = (ctypes.c_size_t * 24)()
            _lib.load().rtda_lovasz_debug_layout(4, 19, 512 * 1024, 0, o)
            from rtda_semanticsegmentation_b200.lovasz_loss import _workspaces
            ws = next(iter(_workspaces.values()))
            S, NBK = int(o[14]), int(o[22])
            cur = ws[int(o[18]):int(o[18]) + 8 * S * NBK].view(torch.int64).view(S, NBK).cpu().numpy()
            occ = cur & 0xffffffff
            nbs = ws[int(o[17]):int(o[17]) + 4 * S].view(torch.int32).cpu().numpy() & 0x7fffffff
            print(f"    buckets per segment {nbs.tolist()}")
            for sgi in (0, 7):
                o_ = occ[sgi, :nbs[sgi]]
                top = np.argsort(-o_)[:6]
                print(f"    seg {sgi}: largest buckets {[(int(i), int(o_[i])) for i in top]}; first 4 {o_[:4].tolist()} last 4 {o_[-4:].tolist()}; "
                      f"buckets > 1.25 x mean: {int((o_ > 1.25 * o_.mean()).sum())}, std/mean {o_.std() / o_.mean():.3f}")
            print(f"    bucket occupancy: max {occ.max()}, mean of non-empty {occ[occ > 0].mean():.0f}, "
                  f"p99 {np.percentile(occ[occ > 0], 99):.0f}, min non-empty {occ[occ > 0].min()}")
        print(f"c2 {path}: {dt * 1e3:.3f} ms/step wall (with profiling events), lsd-segments {int(fb.sum())}")
        for k, (ms, cnt) in prof.items():
            print(f"    {k:26s} {ms / cnt * 1e3:9.1f} us x {cnt // 10}/step")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
