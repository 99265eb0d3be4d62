"""GPU diagnostics: checks every stage of the Lovasz pipeline against numpy on small inputs.
Used while developing the kernels (the build container has no GPU); not part of the product."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import c_oracle  # noqa: E402
from rtda_semanticsegmentation_b200 import _lib, synthetic  # noqa: E402
from rtda_semanticsegmentation_b200.lovasz_loss import _forward_native  # noqa: E402


def layout(b, c, hw, per_image):
    o = (ctypes.c_size_t * 24)()
    _lib.check(_lib.load().rtda_lovasz_debug_layout(b, c, hw, int(per_image), o), "layout")
    names = ["hist", "fghist", "tickets", "desc32", "desc64", "zero_end", "base", "fgbase", "seg_fg", "codes",
             "partial", "buf_a", "buf_b", "total", "S", "T"]
    return dict(zip(names, [int(x) for x in o]))


def stage_check(probas, labels, ignore=255, per_image=False):
    b, c, h, w = probas.shape
    hw = h * w
    L = layout(b, c, hw, per_image)
    p = probas.cuda().contiguous()
    l = labels.cuda().contiguous()
    loss, (unit, codes_t), outs, ws = _forward_native(p, l, ignore, per_image, True)
    torch.cuda.synchronize()
    wsn = ws.cpu().numpy()
    G = b if per_image else 1
    n = b * hw // G
    S, T = L["S"], L["T"]
    pn = probas.numpy().reshape(G, b // G, c, hw)
    ln = labels.numpy().reshape(G, n)
    valid = (ln != ignore) & (ln >= 0) & (ln < c)
    ok = True
    codes = codes_t.cpu().numpy().reshape(G, n)
    want_codes = np.where(valid, ln, 255).astype(np.uint8)
    print("codes ok:", np.array_equal(codes, want_codes)); ok &= np.array_equal(codes, want_codes)
    hist = wsn[L["hist"]:L["hist"] + S * 4 * 256 * 4].view(np.uint32).reshape(S, 4, 256)
    segfg = wsn[L["seg_fg"]:L["seg_fg"] + S * 4].view(np.uint32)
    bufa = wsn[L["buf_a"]:L["buf_a"] + S * n * 8].view(np.uint64).reshape(S, n)
    for g in range(G):
        for cc in range(c):
            s = g * c + cc
            pc = pn[g, :, cc, :].reshape(n)
            fg = valid[g] & (ln[g] == cc)
            err = np.abs(fg.astype(np.float32) - pc).astype(np.float32)
            key = (~err.view(np.uint32)).astype(np.uint32)
            for ps in range(4):
                want = np.bincount((key[valid[g]] >> (8 * ps)) & 255, minlength=256)
                if not np.array_equal(want, hist[s, ps]):
                    print(f"hist mismatch seg {s} pass {ps}: diff at", np.nonzero(want != hist[s, ps])[0][:8]); ok = False
            if segfg[s] != fg.sum():
                print(f"seg_fg mismatch seg {s}: {segfg[s]} vs {fg.sum()}"); ok = False
            if fg.sum() == 0:
                continue
            # buf_a after pass 2 must be the stable sort by the low 24 bits of the key
            pay = (fg.astype(np.uint32) << 31) | (np.arange(n, dtype=np.uint32))
            negb = ((fg.astype(np.float32) - pc) < 0) & valid[g]
            pay |= negb.astype(np.uint32) << 30
            vi = np.nonzero(valid[g])[0]                       # ignored pixels are dropped by pass 0
            order = vi[np.argsort(key[vi] & 0xFFFFFF, kind="stable")]
            want_pairs = (key[order].astype(np.uint64) << 32) | pay[order].astype(np.uint64)
            got_pairs = bufa[s][:len(vi)]
            if not np.array_equal(want_pairs, got_pairs):
                bad = np.nonzero(want_pairs != got_pairs)[0]
                print(f"buf_a mismatch seg {s}: {len(bad)} of {n}, first at {bad[:5]}"); ok = False
    want = c_oracle.lovasz(probas.numpy(), labels.numpy(), ignore, per_image)
    print("loss", float(loss), "oracle", float(want["loss"]))
    print("class_loss max diff", np.abs(outs.class_loss.cpu().numpy() - want["class_loss"]).max())
    print("present ok", np.array_equal(outs.class_present.cpu().numpy(), want["present"]),
          "n_present", outs.n_present.cpu().numpy(), want["n_present"])
    return ok


if __name__ == "__main__":
    torch.manual_seed(0)
    all_ok = True
    for shape, per_image in [((1, 3, 8, 8), False), ((2, 19, 33, 47), False), ((2, 19, 64, 96), True), ((1, 19, 128, 256), False)]:
        b, c, h, w = shape
        g = torch.Generator().manual_seed(h * w)
        pr = torch.softmax(torch.randn(shape, generator=g) * 2, 1)
        lb = torch.randint(0, c, (b, h, w), generator=g)
        lb[torch.rand((b, h, w), generator=g) < 0.1] = 255
        print("==== shape", shape, "per_image", per_image)
        all_ok &= stage_check(pr, lb, 255, per_image)
    print("ALL STAGES OK" if all_ok else "STAGE FAILURES")
