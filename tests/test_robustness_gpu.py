"""GPU: what a memory checker would look for, done by hand (compute-sanitizer is closed on this pool,
profiles/r2_sanitizer_refusal.md), plus the contract edges around the kernels:

* uninitialised reads - every output buffer and the whole workspace pre-filled with 0xFF / NaN bytes must
  give the same bits as zero-filled ones (the forward documents that rows of absent classes and ignored
  pixels of unit_grad are NOT written: the backward must never read them);
* lifetime - a captured CUDA graph keeps replaying correctly after a larger shape replaced the cached workspace;
* NaN / inf probabilities propagate into the loss like the reference's torch.dot;
* batch slices of odd-sized maps (4-byte aligned only), the flat API's gradient, the graph mode of the
  modules, StepMetrics on the device, the `out` checks of the confusion matrix.
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import c_oracle, lovasz_oracle
import rtda_semanticsegmentation_b200 as rtda
from rtda_semanticsegmentation_b200 import _lib, lovasz_loss, synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["bucket", "lsd"])
def sort_path(request):
    rtda.set_sort_path(request.param)
    yield request.param
    rtda.set_sort_path("bucket")


def raw_fwd_bwd(probas, labels, per_image, fill, grad_out=1.0, logits=None, final=False):
    """Calls the C ABI directly with every output buffer and the workspace pre-filled with `fill` bytes.
    Returns the bytes of (loss, class_loss, present, n_present, grad)."""
    lib = _lib.load()
    dev = torch.device("cuda")
    x = (logits if logits is not None else probas).to(dev).contiguous()
    l = labels.to(dev).contiguous()
    b, c, h, w = x.shape
    hw, groups = h * w, (b if per_image else 1)
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.rtda_lovasz_workspace_bytes(b, c, hw, int(per_image), ctypes.byref(nbytes)), "ws")

    def buf(n, dtype):
        t = torch.empty(n, dtype=torch.uint8, device=dev)
        t.fill_(fill)
        return t.view(dtype) if dtype != torch.uint8 else t

    ws = buf(nbytes.value, torch.uint8)
    unit, codes = buf(4 * b * c * hw, torch.float32), buf(b * hw + (-b * hw) % 4, torch.uint8)
    outs = lovasz_loss.LovaszOutputs(buf(lovasz_loss.LovaszOutputs.nbytes(groups, c), torch.uint8), groups, c)
    grad = buf(4 * b * c * hw, torch.float32)
    go = torch.tensor([grad_out], dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    if logits is None and final:
        # final-gradient mode: the forward writes every element of the NCHW gradient; the backward scales in place
        _lib.check(lib.rtda_lovasz_fwd(x.data_ptr(), l.data_ptr(), _lib.label_kind(l.dtype), b, c, hw, 255, 1, int(per_image),
                                       ws.data_ptr(), nbytes.value, _lib.GRAD_FINAL, 1.0, grad.data_ptr(), None, outs.ptr("applied"),
                                       outs.ptr("loss"), outs.ptr("class_loss"), outs.ptr("class_present"), outs.ptr("n_present"),
                                       outs.ptr("group_loss"), outs.ptr("label_errors"), outs.ptr("packed"), st), "fwd")
        _lib.check(lib.rtda_lovasz_bwd_inplace(go.data_ptr(), grad.data_ptr(), outs.ptr("applied"), ws.data_ptr(), b, c, hw, st), "bwd_inplace")
        extra = ()
    elif logits is None:
        _lib.check(lib.rtda_lovasz_fwd(x.data_ptr(), l.data_ptr(), _lib.label_kind(l.dtype), b, c, hw, 255, 1, int(per_image),
                                       ws.data_ptr(), nbytes.value, _lib.GRAD_UNIT, 1.0, unit.data_ptr(), codes.data_ptr(), None, outs.ptr("loss"),
                                       outs.ptr("class_loss"), outs.ptr("class_present"), outs.ptr("n_present"),
                                       outs.ptr("group_loss"), outs.ptr("label_errors"), outs.ptr("packed"), st), "fwd")
        _lib.check(lib.rtda_lovasz_bwd(go.data_ptr(), unit.data_ptr(), codes.data_ptr(), outs.ptr("class_present"),
                                       outs.ptr("n_present"), b, c, hw, int(per_image), grad.data_ptr(), st), "bwd")
        extra = ()
    else:
        pr = buf(4 * b * c * hw, torch.float32)
        _lib.check(lib.rtda_hybrid_fwd(x.data_ptr(), l.data_ptr(), _lib.label_kind(l.dtype), b, c, hw, 255, 1, int(per_image),
                                       ws.data_ptr(), nbytes.value, pr.data_ptr(), unit.data_ptr(), codes.data_ptr(),
                                       outs.ptr("loss"), outs.ptr("loss_ce"), outs.ptr("class_loss"), outs.ptr("class_present"),
                                       outs.ptr("n_present"), outs.ptr("group_loss"), outs.ptr("label_errors"),
                                       outs.ptr("packed"), outs.ptr("n_valid"), st), "hybrid_fwd")
        _lib.check(lib.rtda_hybrid_bwd(go.data_ptr(), go.data_ptr(), pr.data_ptr(), unit.data_ptr(), codes.data_ptr(),
                                       outs.ptr("class_present"), outs.ptr("n_present"), outs.ptr("n_valid"), b, c, hw,
                                       int(per_image), grad.data_ptr(), st), "hybrid_bwd")
        extra = (outs.loss_ce.cpu().numpy().tobytes(), pr.cpu().numpy().tobytes())
    torch.cuda.synchronize()
    return (outs.loss.cpu().numpy().tobytes(), outs.class_loss.cpu().numpy().tobytes(), outs.class_present.cpu().numpy().tobytes(),
            outs.n_present.cpu().numpy().tobytes(), outs.packed.cpu().numpy().tobytes(), grad.cpu().numpy().tobytes()) + extra


CASES = {
    "conf_global": lambda: synthetic.confident_case(2, 64, 192, seed=3) + (False,),       # HW % 4096 == 0: final-gradient mode with 2 images per group
    "conf_per_image_odd": lambda: synthetic.confident_case(3, 33, 47, seed=4) + (True,),
    "absent_classes_and_empty_image": lambda: _absent_case(),
    "saturated_ties": lambda: _saturated_case(),
}


def _absent_case():
    p, l = synthetic.confident_case(3, 40, 64, seed=5)
    l[l == 3] = 7           # class 3 absent everywhere
    l[l == 11] = 255        # class 11 only as ignored
    l[2] = 255              # an image with nothing valid
    return p, l, True


def _saturated_case():
    lab = synthetic.blocky_labels(1, 64, 96, seed=6)
    lg = synthetic.confident_logits(lab, seed=6, scale=3.0, boost=18.0)     # many probabilities exactly 1.0 / 1 - k ulp
    return torch.softmax(lg, 1), lab, False


@pytest.mark.parametrize("final", [True, False], ids=["final_grad", "unit_grad"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_poisoned_buffers_give_identical_bits(case, final, sort_path):
    """Both gradient modes of the forward.  In final-gradient mode the forward itself must write EVERY element
    of the gradient (zeros at ignored pixels, in absent classes, in images without a valid pixel)."""
    probas, labels, per_image = CASES[case]()
    if final and not _lib.load().rtda_lovasz_final_grad_supported(probas.size(0), probas.size(1), probas.size(2) * probas.size(3), int(per_image)):
        pytest.skip("final-gradient mode needs HW % 4096 == 0 with several images per group")
    go = 0.5 if final else 0.75                 # the in-place backward is exact for powers of two
    clean = raw_fwd_bwd(probas, labels, per_image, 0x00, grad_out=go, final=final)
    for fill in (0xFF, 0x7F):                  # 0xFFFFFFFF = NaN, 0x7F7F7F7F = 3.4e38: any read of it shows
        dirty = raw_fwd_bwd(probas, labels, per_image, fill, grad_out=go, final=final)
        for name, a, b in zip(("loss", "class_loss", "present", "n_present", "packed", "grad"), clean, dirty):
            assert a == b, f"{case}: {name} depends on the initial contents of a buffer (fill {fill:#x})"
    want = c_oracle.lovasz(probas.numpy(), labels.numpy(), 255, per_image, grad_out=go)
    assert np.array_equal(np.frombuffer(clean[5], np.float32).reshape(probas.shape), want["grad"])


def test_poisoned_buffers_hybrid_path():
    lab = synthetic.blocky_labels(2, 48, 80, seed=8)
    lab[1, :20] = 255
    lg = synthetic.confident_logits(lab, seed=8)
    clean = raw_fwd_bwd(None, lab, False, 0x00, logits=lg)
    dirty = raw_fwd_bwd(None, lab, False, 0xFF, logits=lg)
    for i, (a, b) in enumerate(zip(clean, dirty)):
        assert a == b, f"hybrid output {i} depends on the initial contents of a buffer"


def test_graph_replay_survives_workspace_replacement():
    """lovasz_loss._workspace: a graph captured on the small shape keeps its workspace alive when a
    larger shape on the same stream replaces the cached one."""
    rtda.lovasz_loss.release_workspaces(captured=True)
    p, l = synthetic.confident_case(1, 64, 96, seed=11, device="cuda")
    want = c_oracle.lovasz(p.cpu().numpy(), l.cpu().numpy(), 255, False)
    crit = rtda.LovaszSoftmax(ignore=255)
    pin = p.clone().requires_grad_(True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        crit(pin, l).backward()
        pin.grad = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            loss = crit(pin, l)
            loss.backward()
        # a much larger shape on the SAME stream: the cache entry of (device, stream) is replaced
        big_p, big_l = synthetic.confident_case(2, 256, 512, seed=12, device="cuda")
        bp = big_p.requires_grad_(True)
        crit(bp, big_l).backward()
        junk = [torch.full((1 << 22,), float("nan"), device="cuda") for _ in range(8)]   # reuse whatever was freed
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    pin.grad.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(want["loss"])) <= 2e-6 * float(want["loss"])
    assert np.array_equal(pin.grad.cpu().numpy(), want["grad"])
    del junk


def test_capture_mode_of_the_modules():
    """LovaszSoftmax(capture=True) / HybridCELovasz(capture=True): same bits as the eager modules, repeated
    calls with different data, a scaled backward."""
    crit_e, crit_g = rtda.LovaszSoftmax(ignore=255), rtda.LovaszSoftmax(ignore=255, capture=True)
    for seed, scale in ((21, 1.0), (22, 0.5), (23, 3.0)):
        p, l = synthetic.confident_case(2, 64, 96, seed=seed, device="cuda")
        a, b = p.clone().requires_grad_(True), p.clone().requires_grad_(True)
        le, lg = crit_e(a, l), crit_g(b, l)
        (le * scale).backward()
        (lg * scale).backward()
        torch.cuda.synchronize()
        assert float(le) == float(lg)
        assert torch.equal(a.grad, b.grad)
    # non-leaf input: the producer's backward consumes the static gradient buffer
    lab = synthetic.blocky_labels(2, 64, 96, seed=24, device="cuda")
    x = synthetic.confident_logits(lab, seed=24)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    crit_e(torch.softmax(xa, 1), lab).backward()
    crit_g(torch.softmax(xb, 1), lab).backward()
    assert torch.equal(xa.grad, xb.grad)
    he, hg = rtda.HybridCELovasz(ignore=255), rtda.HybridCELovasz(ignore=255, capture=True)
    for seed in (25, 26):
        lab = synthetic.blocky_labels(2, 64, 96, seed=seed, device="cuda")
        x = synthetic.confident_logits(lab, seed=seed)
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        ta, tb = he(xa, lab), hg(xb, lab)
        ta.backward()
        tb.backward()
        torch.cuda.synchronize()
        assert float(ta) == float(tb) and float(he.last_ce) == float(hg.last_ce) and float(he.last_lovasz) == float(hg.last_lovasz)
        assert torch.equal(xa.grad, xb.grad)
    # static inputs: no copy at all
    p, l = synthetic.confident_case(2, 64, 96, seed=27, device="cuda")
    sp, sl = crit_g.static_inputs(p, l)
    sp.copy_(p)
    sl.copy_(l)
    spr = sp.requires_grad_(True)
    crit_g(spr, sl).backward()
    a = p.clone().requires_grad_(True)
    crit_e(a, l).backward()
    assert torch.equal(a.grad, spr.grad)


@pytest.mark.parametrize("bad", [float("nan"), float("inf")])
def test_nonfinite_probabilities_reach_the_loss(bad, sort_path):
    """the fixed-point loss sums turn NaN into 0; the reference's torch.dot propagates it."""
    l = synthetic.blocky_labels(1, 48, 64, seed=31, tile=8)
    p = torch.softmax(synthetic.confident_logits(l, seed=31), 1)
    valid = (l[0] != 255).nonzero()
    y, x = [int(v) for v in valid[len(valid) // 2]]
    cls = int(l[0, y, x])
    present = [int(c) for c in torch.unique(l) if int(c) not in (255, cls)]
    other = present[len(present) // 2]
    p[0, other, y, x] = bad                                # a background element of class `other` at a valid pixel
    ref = lovasz_oracle.lovasz_softmax(p, l, ignore=255, stable=True)
    assert not torch.isfinite(ref)
    crit = rtda.LovaszSoftmax(ignore=255)
    loss = crit(p.cuda().requires_grad_(True), l.cuda())
    got = float(loss)
    assert np.isnan(got) if np.isnan(float(ref)) else got == float(ref), (got, float(ref))
    cl = crit.last_outputs.class_loss.cpu().numpy()[0]
    assert not np.isfinite(cl[other]) and np.isfinite(np.delete(cl, other)).all()
    # the same value at an IGNORED pixel takes no part (lovasz_loss.py:157-161)
    p2, l2 = synthetic.confident_case(1, 48, 64, seed=31)
    ig = (l2[0] == 255).nonzero()[0]
    p2[0, 3, int(ig[0]), int(ig[1])] = bad
    assert np.isfinite(float(rtda.LovaszSoftmax(ignore=255)(p2.cuda(), l2.cuda())))


def test_batch_slices_of_odd_sized_maps(sort_path):
    """probas[i:i+1] of a 33x33 map starts at a multiple of C*HW*4 bytes that is not 16-aligned;
    the per_image composition the module docstring defines must work on it."""
    p, l = synthetic.confident_case(3, 33, 33, seed=41, device="cuda")
    crit = rtda.LovaszSoftmax(ignore=255)
    parts = []
    grads = []
    for i in range(3):
        pi = p[i:i + 1].detach().requires_grad_(True)      # contiguous view at an odd offset
        assert pi.data_ptr() % 16 != 0 or i == 0
        li = crit(pi, l[i:i + 1])
        li.backward()
        parts.append(float(li))
        grads.append(pi.grad.cpu().numpy())
    for i in range(3):
        want = c_oracle.lovasz(p[i:i + 1].cpu().numpy(), l[i:i + 1].cpu().numpy(), 255, False)
        assert abs(parts[i] - float(want["loss"])) <= 2e-6 * float(want["loss"])
        assert np.array_equal(grads[i], want["grad"])
    x = torch.randn(3, 19, 33, 33, device="cuda")
    t0 = rtda.hybrid_ce_lovasz(x[1:2].detach().requires_grad_(True), l[1:2])[0]
    t1 = rtda.hybrid_ce_lovasz(x[1:2].clone().requires_grad_(True), l[1:2])[0]
    assert float(t0) == float(t1)
    # a misaligned view of a map whose size IS a multiple of 4 is cloned by the Python layer
    flat = torch.rand(1 + 2 * 4 * 8 * 8, device="cuda")
    pv = torch.softmax(flat[1:].view(2, 4, 8, 8), 1)
    off = torch.empty(1 + pv.numel(), device="cuda")
    off[1:] = pv.reshape(-1)
    view = off[1:].view(2, 4, 8, 8)
    lab = torch.randint(0, 4, (2, 8, 8), device="cuda")
    assert float(rtda.lovasz_softmax(view, lab)) == float(rtda.lovasz_softmax(pv.contiguous(), lab))


def test_flat_api_gradient(sort_path):
    """lovasz_softmax_flat(probas [P,C], labels [P]) (lovasz_loss.py:122-183): loss AND gradient."""
    p4, l4 = synthetic.confident_case(2, 24, 40, seed=51)
    flat, lab = lovasz_oracle.flatten_nchw(p4, l4)
    keep = lab != 255
    fp = flat[keep].clone().requires_grad_(True)            # what Berman's flatten_probas hands over
    want = lovasz_oracle.lovasz_flat(fp, lab[keep], ignore=None, stable=True)
    want.backward()
    gp = flat[keep].clone().cuda().requires_grad_(True)
    got = rtda.lovasz_softmax_flat(gp, lab[keep].cuda(), classes="present", ignore=None)
    got.backward()
    assert float(got) == pytest.approx(float(want), rel=2e-6)
    assert float((gp.grad.cpu() - fp.grad).abs().max()) <= 1e-6 * float(fp.grad.abs().max())
    # through the module method, with the ignore filter inside
    crit = rtda.LovaszSoftmax(ignore=255)
    gp2 = flat.clone().cuda().requires_grad_(True)
    got2 = crit.lovasz_softmax_flat(gp2, lab.cuda())
    got2.backward()
    assert float(got2) == pytest.approx(float(want), rel=2e-6)
    assert not gp2.grad.cpu()[~keep].any()


def test_step_metrics_on_the_device():
    """running sums of the step's losses without a synchronisation per step."""
    dev = torch.device("cuda")
    m = rtda.StepMetrics(("loss_total", "loss_ce", "loss_lovasz"), dev, every=4)
    crit = rtda.HybridCELovasz(ignore=255)
    sums = np.zeros(3)
    last = None
    for step in range(1, 9):
        lab = synthetic.blocky_labels(1, 32, 48, seed=60 + step, device=dev)
        x = synthetic.confident_logits(lab, seed=60 + step).requires_grad_(True)
        total = crit(x, lab)
        total.backward()
        m.update(loss_total=total, loss_ce=crit.last_ce, loss_lovasz=crit.last_lovasz)
        last = np.array([float(total), float(crit.last_ce), float(crit.last_lovasz)])
        sums += last
        r = m.maybe_read(step)
        assert (r is not None) == (step % 4 == 0)
        if r is not None:
            assert r["steps"] == step
            assert [r["loss_total"], r["loss_ce"], r["loss_lovasz"]] == pytest.approx(list(last), rel=1e-7)
            assert [r["loss_total_sum"], r["loss_ce_sum"], r["loss_lovasz_sum"]] == pytest.approx(list(sums), rel=1e-6)
    means = m.means()
    assert means["loss_total"] == pytest.approx(sums[0] / 8, rel=1e-6)
    assert m._buf.is_cuda and m._host.is_pinned()
    m.reset()
    assert m.read()["loss_total_sum"] == 0.0


def test_confusion_matrix_out_is_checked():
    """the kernel does 64-bit atomics on n_class^2 slots of `out`."""
    lab, prd = synthetic.validation_maps(1, 64, 64, seed=71, device="cuda")
    good = torch.zeros(19, 19, dtype=torch.int64, device="cuda")
    rtda.confusion_matrix(lab, prd, 19, 255, out=good)
    for bad in (torch.zeros(19, 19, dtype=torch.int32, device="cuda"),
                torch.zeros(18, 19, dtype=torch.int64, device="cuda"),
                torch.zeros(19, 38, dtype=torch.int64, device="cuda")[:, ::2],
                torch.zeros(19, 19, dtype=torch.int64)):
        with pytest.raises(ValueError, match="out must be"):
            rtda.confusion_matrix(lab, prd, 19, 255, out=bad)
        with pytest.raises(ValueError, match="out must be"):
            rtda.confusion_matrix_from_logits(torch.randn(1, 19, 64, 64, device="cuda"), lab, 255, out=bad)
    assert int(good.sum()) == int((lab != 255).sum())


def test_inplace_backward_on_an_unaligned_gradient_buffer():
    """rtda_lovasz_bwd_inplace takes the scalar path when the gradient buffer is not 16-byte aligned (a view at an odd
    offset of a map whose size is not a multiple of 4); the result is the aligned one."""
    lib = _lib.load()
    probas, labels = synthetic.confident_case(2, 33, 47, seed=81)
    b, c, h, w = probas.shape
    hw = h * w
    dev = torch.device("cuda")
    x, l = probas.to(dev).contiguous(), labels.to(dev).contiguous()
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.rtda_lovasz_workspace_bytes(b, c, hw, 1, ctypes.byref(nbytes)), "ws")
    st = torch.cuda.current_stream().cuda_stream
    go = torch.tensor([2.0], dtype=torch.float32, device=dev)
    got = []
    for off in (0, 1, 3):
        ws = torch.zeros(nbytes.value, dtype=torch.uint8, device=dev)
        flat = torch.full((b * c * hw + 4,), float("nan"), dtype=torch.float32, device=dev)
        grad = flat[off:off + b * c * hw]
        outs = lovasz_loss.LovaszOutputs(torch.zeros(lovasz_loss.LovaszOutputs.nbytes(b, c), dtype=torch.uint8, device=dev), b, c)
        _lib.check(lib.rtda_lovasz_fwd(x.data_ptr(), l.data_ptr(), _lib.label_kind(l.dtype), b, c, hw, 255, 1, 1, ws.data_ptr(), nbytes.value,
                                       _lib.GRAD_FINAL, 1.0, grad.data_ptr(), None, outs.ptr("applied"), outs.ptr("loss"), outs.ptr("class_loss"),
                                       outs.ptr("class_present"), outs.ptr("n_present"), outs.ptr("group_loss"), outs.ptr("label_errors"),
                                       outs.ptr("packed"), st), "fwd")
        _lib.check(lib.rtda_lovasz_bwd_inplace(go.data_ptr(), grad.data_ptr(), outs.ptr("applied"), ws.data_ptr(), b, c, hw, st), "bwd_inplace")
        _lib.check(lib.rtda_lovasz_bwd_inplace(go.data_ptr(), grad.data_ptr(), outs.ptr("applied"), ws.data_ptr(), b, c, hw, st), "bwd_inplace")   # same factor again: nothing happens
        torch.cuda.synchronize()
        assert float(outs._view("applied")[0]) == 2.0
        assert torch.isnan(flat[:off]).all() and torch.isnan(flat[off + b * c * hw:]).all()      # nothing written outside
        got.append(grad.clone().view(b, c, h, w).cpu().numpy())
    want = c_oracle.lovasz(probas.numpy(), labels.numpy(), 255, True, grad_out=2.0)
    for g in got:
        assert np.array_equal(g, want["grad"])


def test_unit_gradient_buffer_at_an_odd_offset_with_a_segment_of_a_multiple_of_4():
    """Global mode over four 33x33 maps: the segment length (4356) is a multiple of 4 although HW is not, so the C ABI
    only asks for a 4-byte aligned unit-gradient buffer - k_part's zero fill must then go element by element, not by
    16-byte bulk copies.  The result equals the aligned one and nothing is written outside the buffer."""
    lib = _lib.load()
    probas, labels = synthetic.confident_case(4, 33, 33, seed=83)
    b, c, h, w = probas.shape
    hw = h * w
    dev = torch.device("cuda")
    x, l = probas.to(dev).contiguous(), labels.to(dev).contiguous()
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.rtda_lovasz_workspace_bytes(b, c, hw, 0, ctypes.byref(nbytes)), "ws")
    st = torch.cuda.current_stream().cuda_stream
    go = torch.tensor([1.0], dtype=torch.float32, device=dev)
    got = []
    for off in (0, 1, 2):
        ws = torch.zeros(nbytes.value, dtype=torch.uint8, device=dev)
        flat = torch.full((b * c * hw + 4,), float("nan"), dtype=torch.float32, device=dev)
        unit = flat[off:off + b * c * hw]
        codes = torch.empty(b * hw, dtype=torch.uint8, device=dev)
        outs = lovasz_loss.LovaszOutputs(torch.zeros(lovasz_loss.LovaszOutputs.nbytes(1, c), dtype=torch.uint8, device=dev), 1, c)
        _lib.check(lib.rtda_lovasz_fwd(x.data_ptr(), l.data_ptr(), _lib.label_kind(l.dtype), b, c, hw, 255, 1, 0, ws.data_ptr(), nbytes.value,
                                       _lib.GRAD_UNIT, 1.0, unit.data_ptr(), codes.data_ptr(), None, outs.ptr("loss"), outs.ptr("class_loss"),
                                       outs.ptr("class_present"), outs.ptr("n_present"), outs.ptr("group_loss"), outs.ptr("label_errors"),
                                       outs.ptr("packed"), st), "fwd")
        grad = torch.empty(b, c, h, w, dtype=torch.float32, device=dev)
        _lib.check(lib.rtda_lovasz_bwd(go.data_ptr(), unit.data_ptr(), codes.data_ptr(), outs.ptr("class_present"), outs.ptr("n_present"),
                                       b, c, hw, 0, grad.data_ptr(), st), "bwd")
        torch.cuda.synchronize()
        assert torch.isnan(flat[:off]).all() and torch.isnan(flat[off + b * c * hw:]).all()      # nothing written outside
        got.append(grad.cpu().numpy())
    want = c_oracle.lovasz(probas.numpy(), labels.numpy(), 255, False)
    for g in got:
        assert np.array_equal(g, want["grad"])
