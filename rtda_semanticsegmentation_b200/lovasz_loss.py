"""Drop-in for the reference's ``losses/lovasz_loss.py`` on B200.

Same names, arguments and error behaviour as the reference module
(arman-rn/RTDA-SemanticSegmentation, losses/lovasz_loss.py):

* ``LovaszSoftmax(reduction="mean", ignore=None)``  (:76-200) - ``forward(probas,
  labels)``, ``prob_flatten`` (:93-120), ``lovasz_softmax_flat`` (:122-183);
* free functions ``lovasz_grad`` (:47-70), ``mean`` (:19-44), ``isnan`` (:14-16);
* the Berman spellings BASELINE.json names, which the reference file was adapted
  from: ``lovasz_softmax``, ``lovasz_softmax_flat``, ``flatten_probas``.

The loss and its gradient are computed by the hand-written sm_100a kernels in
``csrc/lovasz.cu`` through the C ABI of ``include/rtda_lovasz.h``; PyTorch only
owns the device memory, the stream and the autograd hook.  There is no CPU path:
CPU tensors raise.

Documented deviations from the reference:

* an input with no valid pixel returns a 0-dim fp32 zero tensor with a zero
  gradient where the reference returns the Python int 0 (which its own callers
  crash on, train_lovasz.py:135);
* ``per_image=True`` (absent from the reference) is
  ``mean(forward(p[i:i+1], l[i:i+1]) for i in range(B))``;
* inside a group of exactly equal errors the gradient follows a stable order
  (ascending pixel index); the reference's order there is whatever the unstable
  CPU sort produced;
* labels outside ``[0, C)`` that are not ``ignore`` are treated as ignored unless
  ``validate_labels=True``, which raises the reference's ``IndexError`` at the
  cost of one host synchronisation.
"""
from __future__ import annotations

from itertools import filterfalse

import torch
import torch.nn as nn

from . import _lib

__all__ = [
    "LovaszSoftmax",
    "lovasz_softmax",
    "lovasz_softmax_flat",
    "lovasz_grad",
    "flatten_probas",
    "mean",
    "isnan",
    "LovaszOutputs",
    "release_workspaces",
    "set_sort_path",
    "segment_sort_paths",
]


def isnan(x):
    """True where ``x`` is NaN (lovasz_loss.py:14-16)."""
    return x != x


def mean(l, ignore_nan=False, empty=0):
    """Generator-friendly mean with the reference's semantics (lovasz_loss.py:19-44):
    sequential sum in iteration order, one value is returned as is, no value
    returns ``empty`` (or raises ValueError when ``empty == "raise"``)."""
    it = iter(l)
    if ignore_nan:
        it = filterfalse(isnan, it)
    count = 0
    total = None
    for v in it:
        total = v if count == 0 else total + v
        count += 1
    if count == 0:
        if empty == "raise":
            raise ValueError("Empty mean")
        return empty
    return total if count == 1 else total / count


class LovaszOutputs:
    """Side outputs of a forward (device tensors, created lazily as views of one small buffer;
    reading them never synchronises): ``class_loss`` [G,C] fp32, ``class_present`` [G,C] i32,
    ``n_present`` [G] i32, ``group_loss`` [G] fp32, ``label_errors`` [1] i32, ``packed`` [4+2C] f64
    (see include/rtda_lovasz.h; after a sharded forward ``packed`` holds the all-reduced sums)."""

    def __init__(self, buf, groups, classes):
        self._buf, self._g, self._c = buf, groups, classes
        self._work = None    # pending asynchronous all-reduce of `packed` (sharded forward with global_images)
        g, c = groups, classes
        o = 8 * (4 + 2 * c)
        self._off = dict(packed=(0, 4 + 2 * c, torch.float64), loss=(o, 1, torch.float32),
                         label_errors=(o + 4, 1, torch.int32), loss_ce=(o + 8, 1, torch.float32),
                         n_valid=(o + 12, 1, torch.int32), n_present=(o + 16, g, torch.int32),
                         group_loss=(o + 16 + 4 * g, g, torch.float32),
                         class_present=(o + 16 + 8 * g, g * c, torch.int32),
                         class_loss=(o + 16 + 8 * g + 4 * g * c, g * c, torch.float32),
                         applied=(o + 16 + 8 * g + 8 * g * c, 1, torch.float32))

    @staticmethod
    def nbytes(groups, classes):
        return 8 * (4 + 2 * classes) + 16 + 8 * groups + 8 * groups * classes + 8

    def ptr(self, name):
        return self._buf.data_ptr() + self._off[name][0]

    def _view(self, name):
        off, n, dt = self._off[name]
        return self._buf[off:off + n * dt.itemsize].view(dt)

    def _packed(self):
        if self._work is not None:          # make the current stream wait for the collective, once
            self._work.wait()
            self._work = None
        return self._view("packed")

    packed = property(_packed)
    loss = property(lambda self: self._view("loss")[0])
    label_errors = property(lambda self: self._view("label_errors"))
    loss_ce = property(lambda self: self._view("loss_ce")[0])        # hybrid path only
    n_valid = property(lambda self: self._view("n_valid"))            # hybrid path only
    n_present = property(lambda self: self._view("n_present"))
    group_loss = property(lambda self: self._view("group_loss"))
    class_present = property(lambda self: self._view("class_present").view(self._g, self._c))
    class_loss = property(lambda self: self._view("class_loss").view(self._g, self._c))


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} must be a CUDA tensor: rtda_semanticsegmentation_b200 has no CPU path "
            "(the reference's CPU implementation is restated only in oracle/ for testing)")


# Scratch for rtda_lovasz_fwd is several hundred MB (two (key,payload) ping-pong buffers); it is
# kept per (device, stream) instead of being re-allocated every step, which would fragment the
# caching allocator and cost milliseconds of cudaMalloc on large batches.  Work on one stream is
# ordered, so successive calls can share it.  release_workspaces() frees them.
_workspaces = {}
_graph_held = []   # workspaces handed out while a CUDA graph was being captured: the graph replays their address


def release_workspaces(captured: bool = False) -> None:
    """Drops the cached workspaces.  Those a captured CUDA graph points to are kept unless
    ``captured=True`` (only after every such graph has been destroyed)."""
    _workspaces.clear()
    if captured:
        _graph_held.clear()


def _workspace(dev, stream_ptr, nbytes):
    key = (dev.index, stream_ptr)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        # a larger shape arrived: the smaller buffer leaves the cache; it is freed here unless a captured
        # graph holds it (below), in which case it lives on in _graph_held and the graph stays valid
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    if torch.cuda.is_current_stream_capturing() and not any(w is ws for w in _graph_held):
        _graph_held.append(ws)
    return ws


SORT_PATHS = {"bucket": 0, "lsd": 1}


def set_sort_path(path: str) -> None:
    """Selects how the forward sorts (process-wide; a test / benchmark aid): ``"bucket"`` (default) is
    the sample-partitioned bucket sort with a per-segment device-side fallback to the LSD radix sort,
    ``"lsd"`` forces the LSD radix sort everywhere.  Results are bit-identical."""
    _lib.check(_lib.load().rtda_lovasz_set_sort_path(SORT_PATHS[path]), "rtda_lovasz_set_sort_path")


def segment_sort_paths(shape, per_image, stream=None):
    """After a forward of ``shape`` = (B, C, H, W) on the current device/stream: int32 tensor [G, C],
    1 where the (group, class) segment was sorted by the LSD radix path, 0 where the bucket path did it
    (absent classes report 0).  Synchronises; debug / test aid."""
    b, c, h, w = shape
    o = (_lib.ctypes.c_size_t * 24)()
    _lib.check(_lib.load().rtda_lovasz_debug_layout(b, c, h * w, int(per_image), o), "rtda_lovasz_debug_layout")
    dev = torch.device("cuda", torch.cuda.current_device())
    ws = _workspaces[(dev.index, torch.cuda.current_stream(dev).cuda_stream if stream is None else stream)]
    groups = b if per_image else 1
    torch.cuda.synchronize()
    return ws[int(o[16]):int(o[16]) + 4 * groups * c].view(torch.int32).view(groups, c).clone()


def _forward_native(probas: torch.Tensor, labels: torch.Tensor, ignore, per_image: bool, need_grad: bool,
                    exact_backward: bool = False, grad_out_hint: float = 1.0):
    """Runs rtda_lovasz_fwd.  probas [B,C,H,W] fp32 contiguous, labels [B,H,W] contiguous.  Returns
    (loss, saved, outs, ws): ``saved`` is (grad [B,C,H,W], None) in final-gradient mode - the forward has
    written the whole gradient for grad_out = 1 - or (unit [G*C*n], codes) with ``exact_backward``."""
    lib = _lib.load()
    b, c, h, w = probas.shape
    hw = h * w
    groups = b if per_image else 1
    dev = probas.device
    if torch.cuda.current_device() != dev.index:
        torch.cuda.set_device(dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    nbytes = _lib.ctypes.c_size_t(0)
    _lib.check(lib.rtda_lovasz_workspace_bytes(b, c, hw, int(per_image), _lib.ctypes.byref(nbytes)),
               "rtda_lovasz_workspace_bytes")
    ws = _workspace(dev, stream, nbytes.value)
    # final-gradient mode where the shape allows it (per image always; global batches of maps with H*W % 4096 == 0), else the unit gradient
    final = need_grad and not exact_backward and bool(lib.rtda_lovasz_final_grad_supported(b, c, hw, int(per_image)))
    grad = torch.empty((b, c, h, w) if final else (b * c * hw,), dtype=torch.float32, device=dev) if need_grad else None
    codes = torch.empty(b * hw, dtype=torch.uint8, device=dev) if need_grad and not final else None
    outs = LovaszOutputs(torch.empty(LovaszOutputs.nbytes(groups, c), dtype=torch.uint8, device=dev), groups, c)
    _lib.check(lib.rtda_lovasz_fwd(
        probas.data_ptr(), labels.data_ptr(), _lib.label_kind(labels.dtype),
        b, c, hw, 0 if ignore is None else int(ignore), int(ignore is not None), int(per_image),
        ws.data_ptr(), nbytes.value, _lib.GRAD_FINAL if final else _lib.GRAD_UNIT, float(grad_out_hint),
        grad.data_ptr() if grad is not None else None, codes.data_ptr() if codes is not None else None,
        outs.ptr("applied") if final else None,
        outs.ptr("loss"), outs.ptr("class_loss"),
        outs.ptr("class_present"), outs.ptr("n_present"), outs.ptr("group_loss"), outs.ptr("label_errors"),
        outs.ptr("packed"), stream), "rtda_lovasz_fwd")
    return outs.loss, (grad, codes), outs, ws


_pending_reduce = {}   # device index -> LovaszOutputs whose asynchronous all-reduce may still be in flight


class _LovaszSoftmaxFunction(torch.autograd.Function):
    """autograd hook around rtda_lovasz_fwd / rtda_lovasz_bwd.  With ``shard_group`` set
    (per_image over a batch sharded across ranks) the forward adds ONE all-reduce(sum) of the
    packed partials and returns the global mean; the backward needs no collective."""

    @staticmethod
    def forward(ctx, probas, labels, ignore, per_image, holder, shard_group, global_images=None, grad_reduction="sum",
                exact_backward=False, grad_out_hint=1.0):
        need_grad = bool(ctx.needs_input_grad[0])
        hint = float(grad_out_hint)
        if shard_group is not None and global_images is not None:
            # the factor of this rank's slice is known on the host: the forward writes the gradient for it
            import torch.distributed as dist

            grp = None if shard_group is True else shard_group
            hint *= float(probas.size(0)) * (float(dist.get_world_size(grp)) if grad_reduction == "mean" else 1.0) / float(global_images)
        loss, (grad, codes), outs, ws = _forward_native(probas, labels, ignore, per_image, need_grad, exact_backward, hint)
        if holder is not None:
            holder.append(outs)
        ctx.shape = tuple(probas.shape)
        ctx.per_image = per_image
        ctx.shard_scale = None
        ctx.final = need_grad and codes is None          # the forward wrote the finished gradient (no class codes kept)
        if shard_group is not None:
            import torch.distributed as dist

            packed = outs.packed
            grp = None if shard_group is True else shard_group
            # DistributedDataParallel averages the ranks' gradients: "mean" pre-multiplies by the world size
            dscale = float(dist.get_world_size(grp)) if grad_reduction == "mean" else 1.0
            if global_images is None:
                dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=grp)
                loss = (packed[0] / packed[1]).float()
                ctx.shard_scale = (float(probas.size(0)) * dscale / packed[1]).float()   # B_local / B_global, on the device
            else:
                # the global image count is known: nothing on this rank's data path depends on the
                # collective.  The forward returns this rank's SHARE of the global mean (the shares sum
                # to it), the all-reduce runs asynchronously and is waited for only when the reduced
                # statistics are read (outs.packed) - ranks do not meet at a barrier every step.
                loss = (packed[0].clone() / float(global_images)).float()
                prev = _pending_reduce.pop(probas.device.index, None)
                if prev is not None and prev._work is not None:
                    prev._packed()                      # at most one collective in flight per device (stream-side wait only)
                outs._work = dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=grp, async_op=True)
                _pending_reduce[probas.device.index] = outs
                ctx.shard_scale = float(probas.size(0)) * dscale / float(global_images)
        if need_grad:
            if ctx.final:
                ctx.ws = ws                              # the in-place backward uses one counter of this workspace
                ctx.save_for_backward(grad, outs._view("applied"))
            else:
                ctx.save_for_backward(grad, codes, outs.class_present, outs.n_present)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        b, c, h, w = ctx.shape
        lib = _lib.load()
        dev = ctx.saved_tensors[0].device
        go = grad_out.detach().to(device=dev, dtype=torch.float32).reshape(1)
        if ctx.shard_scale is not None:
            go = go * ctx.shard_scale
        if torch.cuda.current_device() != dev.index:
            torch.cuda.set_device(dev)
        if ctx.final:
            # the forward wrote the whole gradient for grad_out = 1: nothing to do for the plain backward(),
            # one in-place multiplication otherwise (exact for power-of-two factors)
            grad, applied = ctx.saved_tensors
            _lib.check(lib.rtda_lovasz_bwd_inplace(
                go.data_ptr(), grad.data_ptr(), applied.data_ptr(), ctx.ws.data_ptr(), b, c, h * w,
                torch.cuda.current_stream(dev).cuda_stream), "rtda_lovasz_bwd_inplace")
            return grad, None, None, None, None, None, None, None, None, None
        unit, codes, present, n_present = ctx.saved_tensors
        grad = torch.empty((b, c, h, w), dtype=torch.float32, device=dev)
        _lib.check(lib.rtda_lovasz_bwd(
            go.data_ptr(), unit.data_ptr(), codes.data_ptr(), present.data_ptr(), n_present.data_ptr(),
            b, c, h * w, int(ctx.per_image), grad.data_ptr(),
            torch.cuda.current_stream(dev).cuda_stream), "rtda_lovasz_bwd")
        return grad, None, None, None, None, None, None, None, None, None


def _as_kernel_inputs(probas: torch.Tensor, labels: torch.Tensor):
    if probas.dim() != 4:
        raise ValueError(f"Expected 4D input, but got {probas.dim()}D")  # lovasz_loss.py:106-107
    _require_cuda(probas, "probas")
    _require_cuda(labels, "labels")
    if probas.dtype == torch.float64:
        raise TypeError("probas must be float32 / float16 / bfloat16 (the reference fails on float64 too)")
    if probas.dtype != torch.float32:
        probas = probas.float()  # the reference promotes through `fg.float() - probas_c` (:171)
    probas = probas.contiguous()  # NCHW; a channels_last tensor is re-laid out once
    if probas.data_ptr() % 16 and (probas.size(2) * probas.size(3)) % 4 == 0:
        probas = probas.clone()   # a contiguous view at an odd offset: the float4 paths need 16-byte alignment
    labels.view(-1)  # same failure as lovasz_loss.py:118 for a non-contiguous target
    if labels.dtype not in (torch.int64, torch.int32, torch.uint8):
        labels = labels.long()
    if labels.numel() != probas.size(0) * probas.size(2) * probas.size(3):
        raise ValueError(f"labels {tuple(labels.shape)} do not match probas {tuple(probas.shape)}")
    return probas, labels


def lovasz_softmax(probas, labels, classes="present", per_image=False, ignore=None,
                   validate_labels=False, return_outputs=False, shard_group=None, global_images=None,
                   grad_reduction="sum", exact_backward=False, grad_out_hint=1.0):
    """Multi-class Lovasz-Softmax loss (Berman signature; what LovaszSoftmax.forward
    computes, lovasz_loss.py:185-200).

    probas [B,C,H,W] class probabilities (softmax output), labels [B,H,W] int.
    ``classes`` accepts 'present' and 'all'; the reference skips foreground-less
    classes unconditionally (:167-168) so both average over present classes.
    ``shard_group`` (a process group, or True for the default group) with ``per_image=True``
    treats ``probas`` as this rank's slice of a batch sharded over the ranks (see dist.py);
    with ``global_images`` (the batch size over all ranks) the collective is asynchronous and the
    returned loss is this rank's SHARE of the global mean (the shares sum to it), not the loss itself.
    The local gradient is scaled by ``B_local / B_global`` so that the SUM of the ranks' parameter
    gradients is the single-device gradient; ``grad_reduction="mean"`` multiplies it by the world size
    for DistributedDataParallel, which averages gradients.

    Backward: the forward's sort kernels write the finished gradient for ``grad_out = 1`` straight into the
    NCHW buffer autograd receives; ``loss.backward()`` then launches one kernel that finds nothing to do.  Any
    other ``grad_out`` (a loss weight, an AMP loss scale) multiplies that buffer in place: bit-identical to
    the reference's fp32 sequence ``(grad_out / G) / K * g`` when the factor is a power of two, within 3 ulp
    of it otherwise.  ``grad_out_hint`` names the ``grad_out`` the caller expects to pass (the weight of this loss in
    the total, e.g. 0.5): the forward then writes the gradient for THAT factor and the backward again has nothing
    to do; a wrong hint only costs the multiplication.  The factor of a sharded batch with ``global_images`` is
    folded in automatically.  ``exact_backward=True`` keeps the unit gradient and runs the separate backward kernel,
    bit-exact for every ``grad_out`` (one more pass over the gradient).
    """
    if classes not in ("present", "all"):
        raise ValueError("classes must be 'present' or 'all'")
    if grad_reduction not in ("sum", "mean"):
        raise ValueError("grad_reduction must be 'sum' or 'mean'")
    probas, labels = _as_kernel_inputs(probas, labels)
    if probas.numel() == 0:
        return probas.sum() * 0.0  # lovasz_loss.py:135-136
    if shard_group is not None and not per_image:
        raise ValueError("only per_image=True shards over ranks (global mode needs one sort over all pixels)")
    holder = []
    loss = _LovaszSoftmaxFunction.apply(probas, labels, ignore, bool(per_image), holder, shard_group, global_images,
                                        grad_reduction, bool(exact_backward), float(grad_out_hint))
    outs = holder[0]
    if validate_labels:
        bad = int(outs.label_errors.item())
        if bad:
            raise IndexError(f"{bad} label value(s) outside [0, {probas.size(1)}) and not the ignore index")
    return (loss, outs) if return_outputs else loss


def flatten_probas(probas, labels, ignore=None):
    """[B,C,H,W] -> [P,C], [B,H,W] -> [P], dropping ``ignore`` pixels (Berman's
    helper; the reference's ``prob_flatten`` :93-120 is the ``ignore=None`` case)."""
    if probas.dim() != 4:
        raise ValueError(f"Expected 4D input, but got {probas.dim()}D")
    c = probas.size(1)
    flat = probas.permute(0, 2, 3, 1).contiguous().view(-1, c)
    lab = labels.view(-1)
    if ignore is None:
        return flat, lab
    keep = lab != ignore
    return flat[keep], lab[keep]


def lovasz_softmax_flat(probas, labels, classes="present", ignore=None):
    """Flat form: probas [P,C], labels [P] (lovasz_loss.py:122-183).  The pixels are
    re-laid out class-major once and run through the same kernels."""
    if probas.dim() != 2:
        raise ValueError(f"Expected 2D [P, C] input, but got {probas.dim()}D")
    if probas.numel() == 0:
        return probas * 0.0  # lovasz_loss.py:135-136
    p4 = probas.t().contiguous().view(1, probas.size(1), probas.size(0), 1)
    return lovasz_softmax(p4, labels.reshape(1, -1, 1), classes=classes, ignore=ignore)


def lovasz_grad(gt_sorted):
    """Gradient of the Lovasz extension w.r.t. sorted errors (lovasz_loss.py:47-70)
    for a 1-D 0/1 tensor already in sorted order.  Device tensor in, device tensor
    out.  The fused kernels never call this (they evaluate the same formula per
    element from ranks); it is kept for API parity."""
    _require_cuda(gt_sorted, "gt_sorted")
    gt = gt_sorted.float()
    p = gt.numel()
    gts = gt.sum()
    inter = gts - torch.cumsum(gt, 0)
    union = gts + torch.cumsum(1.0 - gt, 0)
    jac = 1.0 - inter / union
    if p > 1:
        jac[1:p] = jac[1:p] - jac[0:-1]
    return jac


class LovaszSoftmax(nn.Module):
    """``LovaszSoftmax(ignore=255)(probas, labels)`` as in the reference
    (lovasz_loss.py:76-200; constructed at main.py:299-302 and
    main_adversarial.py:233-237).  Stateless; ``per_image``,
    ``validate_labels`` and ``capture`` (CUDA-graph replay, graphed.py) are additions (see the module docstring)."""

    def __init__(self, reduction="mean", ignore=None, per_image=False, validate_labels=False, capture=False,
                 exact_backward=False, grad_out_hint=1.0):
        super().__init__()
        self.reduction = reduction
        self.ignore = ignore
        self.per_image = per_image
        self.validate_labels = validate_labels
        self.exact_backward = exact_backward
        self.grad_out_hint = grad_out_hint     # the weight this loss gets in the total (see lovasz_softmax)
        self.last_outputs = None
        # capture=True: one CUDA graph of the forward and one of the backward per input shape (graphed.py)
        if capture and validate_labels:
            raise ValueError("validate_labels reads a device counter on the host every step; it cannot be captured")
        self._graphed = None
        if capture:
            from .graphed import GraphedCriterion

            self._graphed = GraphedCriterion(lambda p, l: (lovasz_softmax(p, l, classes="present", per_image=self.per_image,
                                                                          ignore=self.ignore, exact_backward=self.exact_backward,
                                                                          grad_out_hint=self.grad_out_hint),))

    def prob_flatten(self, input, target):
        """[B,C,H,W] -> [B*H*W, C] and [B,H,W] -> [B*H*W] (lovasz_loss.py:93-120)."""
        return flatten_probas(input, target, ignore=None)

    def lovasz_softmax_flat(self, probas, labels, classes="present"):
        """Flat loss over [P,C] / [P] with this module's ignore index (:122-183)."""
        return lovasz_softmax_flat(probas, labels, classes=classes, ignore=self.ignore)

    def forward(self, probas, labels):
        if self._graphed is not None and probas.numel() and not torch.cuda.is_current_stream_capturing():
            probas, labels = _as_kernel_inputs(probas, labels)
            (loss,) = self._graphed(probas, labels)
            self.last_outputs = None          # the side outputs live in the graph's static buffers
            return loss
        loss, outs = lovasz_softmax(probas, labels, classes="present", per_image=self.per_image,
                                    ignore=self.ignore, validate_labels=self.validate_labels, return_outputs=True,
                                    exact_backward=self.exact_backward, grad_out_hint=self.grad_out_hint) if probas.numel() else (lovasz_softmax(probas, labels), None)
        self.last_outputs = outs
        return loss

    def static_inputs(self, probas_like, labels_like):
        """capture=True: the static (probas, labels) buffers of this shape; a producer that writes into
        them (e.g. ``torch.softmax(logits, 1, out=buf)``) saves the per-step input copy."""
        if self._graphed is None:
            raise RuntimeError("static_inputs needs capture=True")
        return self._graphed.static_inputs(*_as_kernel_inputs(probas_like, labels_like))
