"""CUDA-graph mode of the criteria: ``LovaszSoftmax(..., capture=True)`` / ``HybridCELovasz(..., capture=True)``.

A forward of the loss is ~14 kernel launches plus a few small torch ops, the backward one more; on small
batches (1x19x512x1024: 0.22 ms of kernels) the host needs about as long to enqueue them as the GPU to
run them.  With ``capture=True`` the module records, once per input shape, ONE CUDA graph of the forward
and ONE of the backward over static buffers and replays them:

    criterion = LovaszSoftmax(ignore=255, capture=True)
    loss = criterion(probas, labels)     # copy into the static inputs, replay the forward graph
    loss.backward()                      # replay the backward graph; probas receives the static gradient buffer

The kernels are the same, so are the bits.  The two input copies (76 + 8 B/px) are skipped when the caller
writes straight into ``criterion.static_inputs(shape...)``.  The gradient handed to autograd is the graph's
static buffer: it is overwritten by the next backward of the same shape (autograd's own accumulation into
a leaf ``.grad`` copies it, and a producer op's backward consumes it before that).

Everything else about capture is already true of the plain modules: they never synchronise and allocate
only through PyTorch, so a caller can equally capture its WHOLE training step with ``torch.cuda.graph``
(that is what ``bench.py`` does); workspaces used under capture stay alive as long as the graph may run
(``lovasz_loss._graph_held``).
"""
from __future__ import annotations

import torch

__all__ = ["GraphedCriterion"]


class _Replay(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, labels, entry):
        if x.data_ptr() != entry.x.data_ptr():
            entry.x.detach().copy_(x)
        if labels.data_ptr() != entry.labels.data_ptr():
            entry.labels.copy_(labels)
        entry.fwd.replay()
        ctx.entry = entry
        return tuple(o.detach().clone() for o in entry.outs)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *grads):
        e = ctx.entry
        for g, slot in zip(grads, e.grad_outs):
            if g is None:
                slot.zero_()
            else:
                slot.copy_(g.reshape(slot.shape))
        e.bwd.replay()
        return e.grad_in, None, None


class _Entry:
    pass


class GraphedCriterion:
    """Caches one (forward graph, backward graph) pair per input signature of ``fn(x, labels) -> tuple of
    0-dim tensors`` (the first requires grad through x)."""

    def __init__(self, fn):
        self.fn = fn
        self.entries = {}

    @staticmethod
    def _key(x, labels):
        return (tuple(x.shape), x.dtype, tuple(labels.shape), labels.dtype, x.device.index)

    def entry(self, x, labels):
        key = self._key(x, labels)
        e = self.entries.get(key)
        if e is None:
            e = self._capture(x, labels)
            self.entries[key] = e
        return e

    def _capture(self, x, labels):
        dev = x.device
        e = _Entry()
        e.x = torch.empty_like(x, memory_format=torch.contiguous_format).requires_grad_(True)
        e.labels = torch.empty_like(labels, memory_format=torch.contiguous_format)
        e.x.detach().copy_(x)
        e.labels.copy_(labels)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):                          # warm-up outside capture: lazy inits, workspace allocation
                outs = self.fn(e.x, e.labels)
                torch.autograd.grad(outs[0], e.x, torch.ones_like(outs[0]), allow_unused=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        e.fwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(e.fwd):
            e.outs = tuple(self.fn(e.x, e.labels))
        e.grad_outs = tuple(torch.zeros_like(o) for o in e.outs)
        e.bwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(e.bwd, pool=e.fwd.pool()):
            live = [(o, g) for o, g in zip(e.outs, e.grad_outs) if o.requires_grad]
            (e.grad_in,) = torch.autograd.grad([o for o, _ in live], e.x, [g for _, g in live])
        return e

    def __call__(self, x, labels):
        return _Replay.apply(x, labels, self.entry(x, labels))

    def static_inputs(self, x_like, labels_like):
        """The static (x, labels) buffers of this signature: a producer that writes into them saves the
        per-step input copies."""
        e = self.entry(x_like, labels_like)
        return e.x.detach(), e.labels
