"""Sync-free training-step metrics.

The reference reads every loss component with ``.item()`` seven or more times per step
(train_lovasz.py:133-158, :299-330): three running sums, three tqdm postfix values and, every
``PRINT_FREQ_BATCH`` steps (config.py:106), three W&B values.  Each ``.item()`` is a device
synchronisation; once the loss itself costs well under a millisecond they dominate the step.

``StepMetrics`` keeps the running sums and the last values ON THE DEVICE (one small fused update per
step, no synchronisation) and brings them to the host only when asked: ``read()`` is one
device-to-host copy of ``2 * len(names)`` doubles.  ``maybe_read(step)`` does that every
``every`` steps, which is how the reference's periodic logging is served.
"""
from __future__ import annotations

import torch

__all__ = ["StepMetrics"]


class StepMetrics:
    """Running sums / last values of named scalar tensors.

    >>> m = StepMetrics(("loss_total", "loss_ce", "loss_lovasz"), device, every=cfg.PRINT_FREQ_BATCH)
    >>> m.update(loss_total=loss_total, loss_ce=loss_ce, loss_lovasz=loss_lovasz)   # no sync
    >>> stats = m.maybe_read(current_global_step)      # dict every `every` steps, else None
    >>> epoch_means = m.means()                        # train_lovasz.py:160-166 (one sync per epoch)
    """

    def __init__(self, names, device, every=100):
        self.names = tuple(names)
        self.every = int(every)
        k = len(self.names)
        self._buf = torch.zeros(2 * k, dtype=torch.float64, device=device)   # [sums | last values]
        self._host = torch.zeros(2 * k, dtype=torch.float64)
        if self._buf.is_cuda:
            self._host = self._host.pin_memory()
        self.steps = 0

    def update(self, **values):
        """Adds one step's values (0-dim tensors on the metrics' device).  Enqueues two tiny kernels; never
        synchronises."""
        k = len(self.names)
        v = torch.stack([values[n].detach().to(torch.float64).reshape(()) for n in self.names])
        self._buf[:k] += v
        self._buf[k:] = v
        self.steps += 1

    def read(self):
        """{name: last value, name + "_sum": running sum, "steps": n} - one device-to-host copy."""
        k = len(self.names)
        self._host.copy_(self._buf, non_blocking=False)
        out = {n: float(self._host[k + i]) for i, n in enumerate(self.names)}
        out.update({n + "_sum": float(self._host[i]) for i, n in enumerate(self.names)})
        out["steps"] = self.steps
        return out

    def maybe_read(self, step):
        return self.read() if self.every > 0 and step % self.every == 0 else None

    def means(self):
        """Per-step means of the running sums (the reference's epoch averages, train_lovasz.py:160-166)."""
        r = self.read()
        n = max(self.steps, 1)
        return {name: r[name + "_sum"] / n for name in self.names}

    def reset(self):
        self._buf.zero_()
        self.steps = 0
