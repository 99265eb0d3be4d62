#!/usr/bin/env python
"""Benchmark of the Lovasz-Softmax fwd+bwd hot path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4|c5]

One "step" = one forward + backward of the loss over one batch of synthetic input
(`LovaszSoftmax(ignore=255)(probas, labels)`; `loss.backward()`).  Prints ONE JSON line.

* N = 1 (default): workload c2 = BASELINE.json configs[1], 4x19x512x1024, global
  (per_image=False), classes='present', D_conf synthetic distribution.
* N > 1 (torchrun, one rank per GPU): the path shards per image, so every rank runs
  per_image=True on its own 4x19x512x1024 slice of a 4N-image batch and ONE NCCL
  all-reduce of the packed partials gives the global loss (weak scaling).  Before anything is
  timed every rank checks the sharded paths (plain and fused CE + Lovasz) against the oracle on a
  small batch: `parity_check` in the line, non-zero exit when it fails.
* `value`   - Mpix/s (pixels = B*H*W incl. ignored) with inputs resident in HBM; several
  input sets larger than L2 are rotated so no step finds its inputs in L2.
* `e2e`     - the same through the public API from pinned HOST buffers: H2D of probas and
  labels and D2H of the loss inside the timed region.
* `roofline` - algorithmic bytes (160 B/px) / step time against the measured
  HBM peak of MEASURED_PEAKS.json; `kernels` lists every kernel's live CUDA-event time.
* `cpu_baseline` - the oracle port of the reference's CPU path timed on this box's cores;
  `torch_cuda_baseline` - the same torch ops (the reference's Python) with device='cuda'.
* `secondary` - short runs of the other BASELINE configs (c1, c3, c5, the fused CE + Lovasz step,
  the confusion matrix c4), each with its own value and roofline fraction.
* `--impl reference` - the reference arm: the same oracle port (the reference is Python
  and does not travel to the GPU box; oracle/lovasz_oracle.py restates it op for op).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C = 19
ALGO_BYTES_PER_PX = 160.0   # read 19 fp32 probs + 1 int64 label, write 19 fp32 grads
CM_BYTES_PER_PX = 16.0      # int64 label + int64 prediction

WORKLOADS = {
    # name: (images at N=1, H, W, per_image at N=1, description)
    "c1": (1, 512, 1024, False, "configs[0]: 1x19x512x1024 global"),
    "c2": (4, 512, 1024, False, "configs[1]: 4x19x512x1024 global, classes=present"),
    "c3": (16, 720, 1280, True, "configs[2]: 16x19x720x1280 per_image, batch sharded over ranks"),
    "c5": (8, 512, 1024, False, "configs[4]: 8x19x512x1024 (global on 1 GPU; per image, 8/N images per rank on N)"),
    # configs[3] is the confusion-matrix path (its own metric): see bench_confusion()
    "c4": (1, 1024, 2048, False, "configs[3]: validation confusion matrix, 1024x2048 int64 label + prediction maps, images sharded over ranks"),
}
CFG_ID = {"c1": 1, "c2": 2, "c3": 3, "c4": 4, "c5": 5}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons polled through NVML (nvidia_ml_py) every ~2 ms DURING the
    timed region; falls back to one nvidia-smi query when NVML is unavailable."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.sm = []
        self.bits = 0
        self.max_mhz = None
        self.power = []
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def _loop(self, nv, h):
        while not self._stop.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
            except Exception as e:  # keep sampling what we can
                self.err = repr(e)
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._loop, args=(nv, h), daemon=True)
            self.thread.start()
        except Exception as e:
            self.err = repr(e)

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.sm:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
                a, b = [float(x) for x in out.strip().split(",")[:2]]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "note": f"nvml unavailable ({self.err}); one nvidia-smi sample after the region"}
            except Exception as e:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"unavailable: {self.err} / {e!r}"], "samples": 0}
        reasons = sorted(name for bit, name in self.REASONS.items() if self.bits & bit)
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------
# CPU / torch-CUDA baselines: the oracle port of the reference's Python (test infrastructure; these
# legs and the reference arm are the only places bench.py executes anything under oracle/)
def reference_port_step(probas, labels, per_image):
    """One fwd+bwd of the reference's path as restated op for op in oracle/lovasz_oracle.py
    (unstable torch.sort exactly like lovasz_loss.py:173), on whatever device the tensors live."""
    from oracle import lovasz_oracle

    lovasz_oracle.lovasz_softmax_with_grad(probas, labels, ignore=255, per_image=per_image, stable=False)


def time_reference_port(probas_cpu, labels_cpu, per_image, budget_s, steps, warmup):
    """Times the oracle port on the WHOLE batch (the same shape our arm runs).  The first call is cold
    (page-in, thread pool start) and is never counted; the number of further warm-up and timed steps is
    cut to what fits `budget_s` (at least one timed step), judged from the first warm call."""
    torch.set_num_threads(os.cpu_count() or 1)

    def one():
        t0 = time.perf_counter()
        reference_port_step(probas_cpu, labels_cpu, per_image)
        return time.perf_counter() - t0

    cold = one()
    first = one()                                   # warm: sizes the run
    fit = int((budget_s - cold - first) / first)    # further warm calls the budget has room for
    if fit < 1:
        times = [first]                             # budget spent: the warm call is the one timed sample
    else:
        times = [one() for _ in range(min(steps, fit))]
    px = probas_cpu.size(0) * probas_cpu.size(2) * probas_cpu.size(3)
    return px, times, (1 if fit < 1 else 2)


def make_sets(workload, imgs, H, W, rank, dev, logits, distribution="conf", min_bytes=400 << 20, max_sets=6):
    """Several distinct input sets, together larger than L2 (126 MB), rotated every step."""
    from rtda_semanticsegmentation_b200 import synthetic

    set_bytes = imgs * H * W * (4 * C + 8)
    n_sets = max(2, min(max_sets, -(-min_bytes // set_bytes)))
    cfg_id = CFG_ID[workload]
    gen = synthetic.confident_case if distribution == "conf" else synthetic.diffuse_case
    sets = []
    for i in range(n_sets):
        seed = 1234 + 100 * cfg_id + rank + 1000 * i
        if logits:       # the hybrid step starts from logits
            l = synthetic.blocky_labels(imgs, H, W, seed=seed, device=dev)
            p = synthetic.confident_logits(l, seed=seed)
        else:
            p, l = gen(imgs, H, W, seed=seed, device=dev)
        sets.append((p.contiguous(), l.contiguous()))
    return sets, set_bytes


class Stepper:
    """fwd+bwd of one criterion over rotating input sets: CUDA-graph replays of the public call (one graph
    per input set) with eager launches as the fallback.  At N > 1 the graph holds the LOCAL part of the
    sharded step - this rank's share of the global mean, which has exactly the sharded gradient - and
    the one all-reduce of the packed partials is issued asynchronously after every replay (capturing
    the NCCL call itself hung, scripts/nccl_graph_probe.py): the composition
    dist.lovasz_softmax_sharded(global_images=...) / hybrid_ce_lovasz_sharded(ce_reduction="rank_mean")."""

    def __init__(self, crit, sets, dev, world, use_graph=True):
        self.crit, self.sets, self.dev, self.world = crit, sets, dev, world
        self.share = 1.0 / world
        self.pending = None
        self.graphs = None
        if use_graph:
            try:
                self._capture()
            except Exception as e:  # noqa: BLE001
                print(f"[bench] CUDA graph capture failed, timing eager launches: {e!r}", file=sys.stderr)
                self.graphs = None

    def _local(self, x, l):
        loss = self.crit(x, l)
        return loss * self.share if self.world > 1 else loss

    def _capture(self):
        graphs = []
        side = torch.cuda.Stream(self.dev)
        for p, l in self.sets:
            pin = p.detach().requires_grad_(True)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._local(pin, l).backward()
                    pin.grad = None
            torch.cuda.current_stream(self.dev).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                gl = self._local(pin, l)
                gl.backward()
            graphs.append((g, gl, pin, self.crit.last_outputs))
        torch.cuda.synchronize()
        self.graphs = graphs

    def eager_step(self, p, l):
        p = p.detach().requires_grad_(True)
        loss = self._local(p, l)
        loss.backward()
        if self.world > 1:
            self._reduce(self.crit.last_outputs)
        return loss, p.grad

    def _reduce(self, outs):
        import torch.distributed as dist

        if os.environ.get("RTDA_BENCH_SKIP_ALLREDUCE"):  # measurement aid: what the per-step collective costs the local step
            return
        if self.pending is not None:
            self.pending.wait()                      # stream-side wait; at most one collective in flight
        self.pending = dist.all_reduce(outs._view("packed"), op=dist.ReduceOp.SUM, async_op=True)

    def run(self, steps):
        """Enqueues `steps` steps; returns the last step's loss tensor."""
        n = len(self.sets)
        if self.graphs is not None:
            for i in range(steps):
                g, gl, _pin, outs = self.graphs[i % n]
                g.replay()
                if self.world > 1:
                    self._reduce(outs)
            loss = self.graphs[(steps - 1) % n][1]
        else:
            for i in range(steps):
                loss, _ = self.eager_step(*self.sets[i % n])
        if self.pending is not None:
            self.pending.wait()
        return loss


def timed(fn, barrier, dev, world):
    """CUDA-event time of fn() in ms, bracketed by barrier + synchronize, max over ranks."""
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), out


def measure_confusion(rank, dev, world, steps, warmup, batch=1, use_graph=True, barrier=None):
    """configs[3]: one step = `batch` 1024x2048 validation map pairs accumulated into the 19x19 matrix
    in ONE call (validation.py:98-109).  Maps are sharded over ranks; ONE all-reduce of the counts at the
    end, inside the timed region.  16 algorithmic bytes per pixel."""
    import rtda_semanticsegmentation_b200 as rtda
    from rtda_semanticsegmentation_b200 import _lib, synthetic

    _, H, W, _, _ = WORKLOADS["c4"]
    n_sets = max(2, -(-10 // batch))                         # >= 10 maps = 335 MB > L2
    maps = [synthetic.validation_maps(batch, H, W, seed=1234 + 400 + rank + 100 * i, device=dev) for i in range(n_sets)]
    px = batch * H * W
    cm = rtda.ConfusionMatrix(19, 255, device=dev)
    for i in range(max(warmup, 3)):
        cm.update(*maps[i % n_sets])
    cm.reset()
    graph = None
    if use_graph:
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                cm.update(*maps[0])
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for l, p in maps:
                    cm.update(l, p)
            cm.reset()
        except Exception as e:  # noqa: BLE001
            print(f"[bench] CUDA graph capture failed, timing eager launches: {e!r}", file=sys.stderr)
            graph = None
    steps = steps if graph is None else max(n_sets, steps // n_sets * n_sets)

    def region():
        if graph is not None:
            for _ in range(steps // n_sets):
                graph.replay()
        else:
            for i in range(steps):
                cm.update(*maps[i % n_sets])
        cm.all_reduce()

    ms, _ = timed(region, barrier, dev, world)
    ms_step = ms / steps
    total = int(cm.compute().sum())
    # per-launch kernel time from the library's CUDA events, eager launches
    _lib.profile_enable(True)
    _lib.profile_read()
    for i in range(max(steps, n_sets)):
        cm.update(*maps[i % n_sets])
    stages = _lib.profile_read()
    _lib.profile_enable(False)
    k_ms, k_n = stages.get("k_confusion", (0.0, 1))
    per = k_ms / max(k_n, 1)
    peak, peak_src = peaks()
    achieved = CM_BYTES_PER_PX * px / (per * 1e-3) / 1e9 if per > 0 else 0.0
    # calibration: what a plain device copy moving the SAME number of bytes reaches at this size (the measured
    # peak is a 2 GB transfer; a 33 MB one pays launch, ramp-up and tail like any kernel of a few microseconds)
    # (one destination per source: a single 17 MB destination would stay in L2 and the copy would move half its bytes)
    src = [torch.empty(px, dtype=torch.int64, device=dev) for _ in range(n_sets)]
    dst = [torch.empty(px, dtype=torch.int64, device=dev) for _ in range(n_sets)]
    for a, d in zip(src, dst):
        d.copy_(a)
    cal_ms_graph = None
    try:
        gcal = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gcal):
            for a, d in zip(src, dst):
                d.copy_(a)
        ms_cal, _ = timed(lambda: [gcal.replay() for _ in range(4)], barrier, dev, world)
        cal_ms_graph = ms_cal / (4 * n_sets)
    except Exception:  # noqa: BLE001
        pass
    del src, dst
    return dict(copy_same_bytes_ms=cal_ms_graph,
                copy_same_bytes_frac=(CM_BYTES_PER_PX * px / (cal_ms_graph * 1e-3) / 1e9 / peak) if cal_ms_graph else None,
                value=world * px / (ms_step * 1e-3) / 1e6, ms_per_step=ms_step, steps=steps, maps_per_step=batch, px=px,
                launch="cuda_graph replay" if graph is not None else "eager", counted_pixels=total, n_sets=n_sets,
                kernel_ms=per, kernel_launches=int(k_n), achieved=achieved, frac=achieved / peak, peak=peak,
                peak_src=peak_src, maps=maps)


def bench_confusion(args, rank, local_rank, world):
    """--workload c4: the confusion-matrix path as the primary metric of the line."""
    import numpy as np

    import rtda_semanticsegmentation_b200 as rtda

    _, H, W, _, desc = WORKLOADS["c4"]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    m = measure_confusion(rank, dev, world, args.steps, args.warmup, batch=args.maps_per_call, use_graph=not args.no_graph, barrier=barrier)
    clocks = sampler.stop() if rank == 0 else None
    maps = m.pop("maps")
    # end to end through the reference signature: numpy maps on the host -> fast_hist -> numpy matrix
    host = [(l[0].cpu().numpy().ravel(), p[0].cpu().numpy().ravel()) for l, p in maps[:3]]
    for i in range(3):
        rtda.fast_hist(*host[i % len(host)], 19)
    barrier()
    t0 = time.perf_counter()
    acc = np.zeros((19, 19))
    for i in range(args.steps):
        acc += rtda.fast_hist(*host[i % len(host)], 19)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    px1 = H * W
    line = {
        "metric": "confusion_matrix_mpix_per_s", "value": m["value"], "unit": "Mpix/s", "n_gpus": world, "steps": m["steps"],
        "warmup": max(args.warmup, 3), "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": f"c4: {desc}", "classes": 19, "ignore": 255, "maps_per_step": m["maps_per_step"],
                   "launch": m["launch"],
                   "l2_policy": f"{m['n_sets']} distinct inputs ({m['n_sets'] * m['px'] * 16 >> 20} MiB > 126 MiB L2) rotated every step",
                   "counted_pixels": m["counted_pixels"]},
        "clocks": clocks,
        "e2e": {"value": world * px1 / (e2e_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": px1 * 16, "d2h_bytes_per_step": 19 * 19 * 8,
                "note": "rtda.fast_hist(label_true_np, label_pred_np, 19): pageable numpy maps -> H2D -> kernel -> 19x19 int64 D2H, wall clock"},
        "gpu_launches": m["kernel_launches"],
        # one launch per step: the step time of the timed region (CUDA events around the replayed graph of back-to-back
        # calls, launch gaps included) IS the kernel's average launch duration, as for the Lovasz line; the per-launch
        # events of eager launches are kept beside it (they add ~3 us to a 10 us kernel, nothing to a 45 us one)
        "roofline": {"bound": "hbm", "achieved": CM_BYTES_PER_PX * m["px"] / (m["ms_per_step"] * 1e-3) / 1e9, "peak": m["peak"], "unit": "GB/s",
                     "frac": CM_BYTES_PER_PX * m["px"] / (m["ms_per_step"] * 1e-3) / 1e9 / m["peak"],
                     "traffic": None, "peak_source": m["peak_src"], "dominant_kernel": "k_confusion",
                     "ms_per_launch": m["ms_per_step"], "what": "16 B/px (int64 label + int64 prediction) x pixels per call / step time of the timed region (one k_confusion launch per step, " + m["launch"] + ")",
                     "eager_events": {"ms_per_launch": m["kernel_ms"], "achieved": m["achieved"], "frac": m["frac"],
                                      "what": "two CUDA events around every eager launch (library profile mode)"},
                     "plain_copy_of_the_same_bytes": {"ms": m["copy_same_bytes_ms"], "frac_of_peak": m["copy_same_bytes_frac"],
                                                      "what": "torch copy_ of the same number of bytes (read + write), graph-replayed back to back: the ceiling a kernel of this size can reach"}},
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle import hist_oracle
        l, p = maps[0][0][0].cpu().numpy(), maps[0][1][0].cpu().numpy()
        hist_oracle.validation_confusion(l, p, 19, 255)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            hist_oracle.validation_confusion(l, p, 19, 255)
            ts.append(time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": px1 / min(ts) / 1e6, "unit": "Mpix/s", "cores": 1, "kind": "port",
                                "sample": "one 1024x2048 map, best of 5; oracle/hist_oracle.py numpy restatement of "
                                          "validation.py:98-109 + utils.py:62-89 (np.bincount is single-threaded)"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (its torch ops, restated in
    oracle/), all host threads, on OUR arm's workload - the whole batch, never a slice of it."""
    if rank != 0:
        return 0
    from rtda_semanticsegmentation_b200 import synthetic

    imgs, H, W, per_image_n1, desc = WORKLOADS[args.workload]
    if args.workload == "c4":
        from oracle import hist_oracle
        lab, prd = synthetic.validation_maps(1, H, W, seed=1234 + 400)
        l, p = lab.numpy(), prd.numpy()
        ts = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            hist_oracle.validation_confusion(l, p, 19, 255)
            if i >= args.warmup:
                ts.append(time.perf_counter() - t0)
        t = sum(ts) / len(ts)
        val = H * W / t / 1e6
        print(json.dumps({
            "impl": "reference", "metric": "confusion_matrix_mpix_per_s", "value": val, "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": {"workload": f"c4: {desc}"},
            "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": 1, "kind": "port",
                             "sample": "one 1024x2048 map per step; oracle/hist_oracle.py numpy restatement (np.bincount is single-threaded)"},
            "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0
    per_image = per_image_n1 or world > 1
    if args.workload == "c3":
        imgs = max(1, imgs // world)
    gen = synthetic.confident_case if args.distribution == "conf" else synthetic.diffuse_case
    probas, labels = gen(imgs, H, W, seed=1234 + 100 * CFG_ID[args.workload])
    px, times, warm_done = time_reference_port(probas, labels, per_image, args.reference_budget_s, args.steps, args.warmup)
    t = sum(times) / len(times)
    val = px / t / 1e6
    line = {
        "impl": "reference", "metric": "lovasz_softmax_fwd_bwd_mpix_per_s", "value": val, "unit": "Mpix/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": warm_done, "steps_requested": args.steps,
        "warmup_requested": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "classes": C, "ignore": 255, "per_image": per_image,
                   "distribution": "D_" + args.distribution, "labels": "int64", "images": imgs, "shape": [imgs, C, H, W]},
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"the whole workload per step ({imgs}x{C}x{H}x{W}); {len(times)} timed step(s) of the "
                                   f"{args.steps} requested fit the {args.reference_budget_s:.0f} s budget (a step takes {t:.1f} s); "
                                   "oracle/lovasz_oracle.py torch-CPU restatement of losses/lovasz_loss.py"},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def parity_check(rank, world, dev):
    """N > 1, before anything is timed: the sharded paths on a small batch against the oracle.  Every rank
    checks its own slice (gradient) and the reduced losses; returns (ok, details)."""
    import numpy as np
    import torch.distributed as dist

    from oracle import c_oracle, lovasz_oracle
    from rtda_semanticsegmentation_b200 import synthetic
    from rtda_semanticsegmentation_b200.dist import hybrid_ce_lovasz_sharded, lovasz_softmax_sharded, shard_range

    b, H, W = 2 * world, 48, 64
    labels = synthetic.blocky_labels(b, H, W, seed=4242, tile=8)
    logits = synthetic.confident_logits(labels, seed=4242)
    probas = torch.softmax(logits, 1)
    lo, hi = shard_range(b, rank, world)
    det = {}
    ok = True
    # plain per-image Lovasz, both forms of the collective
    want = c_oracle.lovasz(probas.numpy(), labels.numpy(), 255, per_image=True)
    gmax = float(np.abs(want["grad"]).max())
    for mode, gi in (("in_stream", None), ("async", b)):
        p = probas[lo:hi].to(dev).requires_grad_(True)
        loss, stats = lovasz_softmax_sharded(p, labels[lo:hi].to(dev), ignore=255, global_images=gi)
        loss.backward()
        got_loss = float(stats["loss"])
        g = p.grad.cpu().numpy()
        d = float(np.abs(g - want["grad"][lo:hi]).max())
        det[f"lovasz_{mode}"] = {"loss_rel": abs(got_loss - float(want["loss"])) / abs(float(want["loss"])),
                                 "grad_maxdiff_over_gmax": d / gmax,
                                 "grad_bits_equal": bool(np.array_equal(g, want["grad"][lo:hi]))}
        ok &= det[f"lovasz_{mode}"]["loss_rel"] <= 2e-6 and d <= 1e-6 * gmax
    # fused CE + 0.5 Lovasz from logits against the reference composition on the whole batch (train_lovasz.py:98-105)
    lg = logits.clone().requires_grad_(True)
    ce = torch.nn.functional.cross_entropy(lg, labels, ignore_index=255)
    lv = lovasz_oracle.lovasz_softmax(torch.softmax(lg, 1), labels, ignore=255, per_image=True, stable=True)
    (ce + 0.5 * lv).backward()
    wg = lg.grad[lo:hi]
    x = logits[lo:hi].to(dev).requires_grad_(True)
    total, gce, glv, _ = hybrid_ce_lovasz_sharded(x, labels[lo:hi].to(dev), ignore=255, lovasz_weight=0.5)
    total.backward()
    d = float((x.grad.cpu() - wg).abs().max()) / float(wg.abs().max())
    det["hybrid_global"] = {"ce_rel": abs(float(gce) - float(ce)) / abs(float(ce)), "lovasz_rel": abs(float(glv) - float(lv)) / abs(float(lv)),
                            "grad_maxdiff_over_gmax": d}
    ok &= det["hybrid_global"]["ce_rel"] <= 1e-5 and det["hybrid_global"]["lovasz_rel"] <= 1e-5 and d <= 1e-5
    # rank_mean form: the shares sum to mean-of-rank-means CE + 0.5 global Lovasz
    x2 = logits[lo:hi].to(dev).requires_grad_(True)
    t2, ce2, lv2, _ = hybrid_ce_lovasz_sharded(x2, labels[lo:hi].to(dev), ignore=255, lovasz_weight=0.5, global_images=b,
                                               ce_reduction="rank_mean")
    t2.backward()
    s = torch.stack([ce2.detach().double(), lv2.detach().double()])
    dist.all_reduce(s)
    ce_rm = sum(float(torch.nn.functional.cross_entropy(logits[a:z], labels[a:z], ignore_index=255))
                for a, z in (shard_range(b, r, world) for r in range(world))) / world
    det["hybrid_rank_mean"] = {"ce_rel": abs(float(s[0]) - ce_rm) / abs(ce_rm), "lovasz_rel": abs(float(s[1]) - float(lv)) / abs(float(lv))}
    ok &= det["hybrid_rank_mean"]["ce_rel"] <= 1e-5 and det["hybrid_rank_mean"]["lovasz_rel"] <= 1e-5
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    return bool(flag.item()), det


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--distribution", default="conf", choices=["conf", "diffuse"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short runs of the other configs")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--hybrid", default="none", choices=["none", "fused", "unfused"],
                    help="time CE + 0.5*Lovasz from LOGITS (train_lovasz.py:98-108): fused = rtda_hybrid_fwd/bwd, "
                         "unfused = torch softmax + torch CE + LovaszSoftmax; not the BASELINE metric")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    ap.add_argument("--reference-budget-s", type=float, default=150.0)
    ap.add_argument("--maps-per-call", type=int, default=1, help="c4: validation maps accumulated by one call")
    ap.add_argument("--sort-path", default="bucket", choices=["bucket", "lsd"],
                    help="forward sort path (A/B aid): sample-partitioned bucket sort (default) or LSD radix sort everywhere")
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return reference_arm(args, rank, world)
    if args.workload == "c4":
        return bench_confusion(args, rank, local_rank, world)
    imgs, H, W, per_image_n1, desc = WORKLOADS[args.workload]

    # ------------------------------------------------------------------ our arm
    import rtda_semanticsegmentation_b200 as rtda
    from rtda_semanticsegmentation_b200 import _lib

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    rtda.set_sort_path(args.sort_path)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = None
    if world > 1:
        ok, parity = parity_check(rank, world, dev)
        if not ok:
            if rank == 0:
                print(json.dumps({"parity_check": "FAILED", "details": parity}))
            dist.destroy_process_group()
            return 3

    per_image = per_image_n1 or world > 1
    if args.workload in ("c3", "c5"):
        imgs = max(1, imgs // world)       # configs[2] / [4]: the batch is split over the ranks
        scaling = "strong"
    else:
        scaling = "weak"

    class Unfused(torch.nn.Module):
        def __init__(self, lov):
            super().__init__()
            self.lov = lov

        def forward(self, x, l):
            out = torch.nn.functional.cross_entropy(x, l, ignore_index=255) + 0.5 * self.lov(torch.softmax(x, dim=1), l)
            self.last_outputs = self.lov.last_outputs
            return out

    def criterion(kind, per_img):
        if kind == "fused":
            return rtda.HybridCELovasz(ignore=255, lovasz_weight=0.5, per_image=per_img)
        # N > 1: the step's loss is this rank's share 1/world of the global mean; the forward is told so and writes the
        # gradient for that factor (what dist.lovasz_softmax_sharded(global_images=...) does), the backward is a no-op
        lov = rtda.LovaszSoftmax(ignore=255, per_image=per_img, grad_out_hint=(1.0 / world) if (world > 1 and kind == "none") else 1.0)
        return Unfused(lov) if kind == "unfused" else lov

    sets, set_bytes = make_sets(args.workload, imgs, H, W, rank, dev, args.hybrid != "none", args.distribution)
    n_sets = len(sets)
    px_step = imgs * H * W
    crit = criterion(args.hybrid, per_image)
    stepper = Stepper(crit, sets, dev, world, use_graph=False)
    for i in range(warmup):
        stepper.eager_step(*sets[i % n_sets])
    barrier()
    stepper = Stepper(crit, sets, dev, world, use_graph=not args.no_graph)
    graphs_on = stepper.graphs is not None

    # ---- timed region 1: device-resident inputs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    host_t = [0.0]

    def region1():
        t0 = time.perf_counter()
        out = stepper.run(args.steps)
        host_t[0] = (time.perf_counter() - t0) * 1e3 / args.steps      # host enqueue time, no sync inside
        return out

    ms_total, loss = timed(region1, barrier, dev, world)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * px_step / (ms_step * 1e-3) / 1e6
    loss_value = float(loss.item())

    # ---- timed region 1a: the same public call launched eagerly (no graph, no profiling events): what a
    # caller of criterion(probas, labels); loss.backward() gets without capturing anything
    eager = Stepper(crit, sets, dev, world, use_graph=False)
    eager.run(3)                                   # untimed: the allocator's blocks of the eager path exist again after the captures
    ms_eager, _ = timed(lambda: eager.run(args.steps), barrier, dev, world)
    eager_ms_step = ms_eager / args.steps

    # ---- region 1b: eager launches with the library's per-launch CUDA events on (kernel table of the roofline)
    _lib.profile_enable(True)
    _lib.profile_read()
    ms_prof, _ = timed(lambda: eager.run(args.steps), barrier, dev, world)
    prof_ms_step = ms_prof / args.steps
    stages = _lib.profile_read()
    _lib.profile_enable(False)

    # ---- timed region 2: end to end from pinned host buffers.  Every step's inputs are copied host -> device
    # inside the region (double-buffered on a side stream so the copy of step i+1 overlaps the kernels of
    # step i) and every step's loss is copied back to the host.
    from rtda_semanticsegmentation_b200.host_pipeline import DeviceStager

    hp = [(p.cpu().pin_memory(), l.cpu().pin_memory()) for p, l in sets[:2]]
    stager = DeviceStager(sets[0], dev, depth=2)
    host_loss = torch.empty((args.steps + 8,), dtype=torch.float32).pin_memory()

    def e2e_run(n):
        slot = stager.submit(hp[0])
        for i in range(n):
            nxt = stager.submit(hp[(i + 1) % 2]) if i + 1 < n else None
            dp, dl = stager.acquire(slot)
            loss, _ = eager.eager_step(dp, dl)
            stager.release(slot)
            host_loss[i].copy_(loss.detach(), non_blocking=True)
            slot = nxt

    e2e_run(3)
    ms_e2e, _ = timed(lambda: e2e_run(args.steps), barrier, dev, world)
    e2e_ms = ms_e2e / args.steps
    e2e_value = world * px_step / (e2e_ms * 1e-3) / 1e6
    assert abs(float(host_loss[args.steps - 1]) - float(eager.eager_step(*sets[(args.steps - 1) % 2])[0])) < 1e-6

    # ---- secondary: short runs of the other configs (every rank takes part; rank 0 reports)
    secondary = {}
    fb_segments = int(rtda.segment_sort_paths((imgs, C, H, W), per_image).sum())
    peak, peak_src = peaks()
    if not args.no_secondary and args.hybrid == "none" and args.workload == "c2" and args.distribution == "conf":
        del stepper, eager, stager
        sec_steps = max(6, min(args.steps, 20))

        def short(name, wl, kind, per_img, images, sc, distribution="conf"):
            """value / roofline fraction of one more config, graph-replayed (eager at capture failure)."""
            try:
                _, h2, w2, _, d2 = WORKLOADS[wl]
                ss, sb = make_sets(wl, images, h2, w2, rank, dev, kind != "none", distribution, min_bytes=300 << 20, max_sets=4)
                cr = criterion(kind, per_img)
                st = Stepper(cr, ss, dev, world, use_graph=False)
                for i in range(3):
                    st.eager_step(*ss[i % len(ss)])
                barrier()
                st = Stepper(cr, ss, dev, world, use_graph=not args.no_graph)
                ms, _ = timed(lambda: st.run(sec_steps), barrier, dev, world)
                per = ms / sec_steps
                px = images * h2 * w2
                # fused CE + Lovasz from logits reads logits + labels and writes grad_logits: the same 160 B/px
                ach = ALGO_BYTES_PER_PX * px / (per * 1e-3) / 1e9
                secondary[name] = {"workload": f"{wl}: {d2}", "distribution": "D_" + distribution, "images_per_rank": images, "per_image": per_img, "loss": kind if kind != "none" else "lovasz",
                                   "value": world * px / (per * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": per, "steps": sec_steps,
                                   "scaling": sc, "launch": "cuda_graph replay" if st.graphs is not None else "eager",
                                   "roofline_frac": ach / peak, "achieved_gbs_per_gpu": ach}
                del st, ss, cr
                rtda.lovasz_loss.release_workspaces()
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                secondary[name] = {"error": repr(e)[:300]}

        if world == 1:
            short("c1", "c1", "none", False, 1, "weak")
            short("c2_per_image", "c2", "none", True, 4, "weak")        # what each rank runs at N > 1
            short("c2_diffuse", "c2", "none", False, 4, "weak", "diffuse")   # D_diffuse: the exact tail pruning skips most buckets
            short("c2_hybrid_fused", "c2", "fused", False, 4, "weak")
            short("c5_global", "c5", "none", False, 8, "strong")
            short("c5_hybrid_fused_global", "c5", "fused", False, 8, "strong")
            short("c3_per_image", "c3", "none", True, 16, "strong")
            short("c3_hybrid_fused_per_image", "c3", "fused", True, 16, "strong")
        else:
            short("c2_hybrid_fused_per_image", "c2", "fused", True, 4, "weak")
            if 16 % world == 0:
                short("c3_per_image", "c3", "none", True, 16 // world, "strong")
                short("c3_hybrid_fused_per_image", "c3", "fused", True, 16 // world, "strong")
            if 8 % world == 0:
                short("c5_hybrid_fused_per_image", "c5", "fused", True, 8 // world, "strong")
        for batch in (1, 8):
            try:
                m = measure_confusion(rank, dev, world, 20, 3, batch=batch, use_graph=not args.no_graph, barrier=barrier)
                m.pop("maps")
                secondary[f"c4_confusion_{batch}map"] = {
                    "workload": "c4: " + WORKLOADS["c4"][4], "maps_per_call": batch, "value": m["value"], "unit": "Mpix/s",
                    "ms_per_step": m["ms_per_step"], "launch": m["launch"], "kernel_ms_per_launch": m["kernel_ms"],
                    "roofline_frac": CM_BYTES_PER_PX * m["px"] / (m["ms_per_step"] * 1e-3) / 1e9 / m["peak"],
                    "achieved_gbs_per_gpu": CM_BYTES_PER_PX * m["px"] / (m["ms_per_step"] * 1e-3) / 1e9, "scaling": "weak",
                    "roofline_frac_eager_events": m["frac"],
                    "plain_copy_of_the_same_bytes": {"ms": m["copy_same_bytes_ms"], "frac_of_peak": m["copy_same_bytes_frac"]}}
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                secondary[f"c4_confusion_{batch}map"] = {"error": repr(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(args.workload)
    except OSError:
        pass
    algo_bytes = ALGO_BYTES_PER_PX * px_step
    achieved = algo_bytes / (ms_step * 1e-3) / 1e9
    # per-kernel compulsory traffic of each stage's own job (bytes per launch), for the kernel table
    E = C * px_step
    # plain Lovasz: the forward writes the finished gradient where the shape allows it (always per image; global batches
    # with HW % 4096 == 0), the backward is then an in-place rescale that exits at once for grad_out = 1
    final_mode = args.hybrid == "none" and bool(_lib.load().rtda_lovasz_final_grad_supported(imgs, C, H * W, int(per_image)))
    own_bytes = {
        "memset_workspace": None,
        "k_prep": px_step * (8 + 1) if args.hybrid != "fused" else E * 8 + px_step * 9,   # labels in, class codes out (+ logits in, probabilities out)
        "k_split": E // 16 * 8,                            # 1/16 sample, half of every 32-byte sector used
        "k_part": E * 4 + E * 1 + E * 8 + E * 4,           # probabilities + codes (re-read per class) in, pairs out, the gradient's zero fill out
        "k_plan": None,
        "k_local": E * 8 + E * 4 // 3,                     # pairs in, the non-zero third of the gradient scattered out
        "k_hist_fb": None,
        "k_scan": None,
        "k_sort_pass<0,probas>": E * 4 + px_step * 1 + E * 8,
        "k_sort_pass<1>": E * 16,
        "k_sort_pass<2>": E * 16,
        "k_rank_last": E * 8 + E * 4,
        "k_finalize": None,
        "k_bwd": (None if final_mode else E * 8) if args.hybrid != "fused" else E * 12 + px_step,
    }
    if fb_segments == 0:
        # no segment fell back: the LSD kernels were launched and exited at once, they moved no data
        for name in ("k_sort_pass<0,probas>", "k_sort_pass<1>", "k_sort_pass<2>", "k_rank_last"):
            own_bytes[name] = None
    kernels = []
    launches = 0
    for name, (ms, cnt) in stages.items():
        if name == "gaps_between_launches":
            continue
        if name != "memset_workspace":
            launches += cnt
        per = ms / max(cnt, 1)
        ob = own_bytes.get(name)
        kernels.append({"name": name, "launches": int(cnt), "ms_per_launch": round(per, 5),
                        "share_of_step": round(ms / args.steps / prof_ms_step, 4),
                        "own_gbs": round(ob / (per * 1e-3) / 1e9, 1) if ob and per > 0 else None,
                        "own_frac_of_peak": round(ob / (per * 1e-3) / 1e9 / peak, 4) if ob and per > 0 else None})
    kernels.sort(key=lambda k: -k["share_of_step"])
    dom = kernels[0]
    gaps = stages.get("gaps_between_launches", (0.0, 0))
    # DRAM floor of the two-movement design: probabilities 4 + codes 1 in, pairs 8 out, gradient zeros 4 out (k_part);
    # pairs 8 in (k_local; its scatter of the non-zero gradients lands on lines the fill wrote) = 25 B per element;
    # with the unit gradient (fused CE + Lovasz, exact_backward) the backward pass reads 4 and writes 4 more = 33; + labels
    floor_per_elem = 25.0 if final_mode else 33.0
    floor_bytes = floor_per_elem * E + 9.0 * px_step
    floor_ms = floor_bytes / (peak * 1e9) * 1e3

    line = {
        "metric": "lovasz_softmax_fwd_bwd_mpix_per_s" if args.hybrid == "none" else f"hybrid_ce_lovasz_fwd_bwd_mpix_per_s[{args.hybrid}]",
        "value": value, "unit": "Mpix/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step, "host_enqueue_ms_per_step": host_t[0],
        "eager_ms_per_step_no_events": eager_ms_step, "eager_ms_per_step_with_profiling_events": prof_ms_step,
        "eager_gap_ms_per_step": gaps[0] / args.steps, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"{args.workload}: {desc}" + ("" if world == 1 else f"; per rank {imgs}x{C}x{H}x{W}, per_image=True sharded over {world} ranks, 1 all-reduce of {4 + 2 * C} doubles"),
            "classes": C, "ignore": 255, "per_image": per_image, "distribution": "D_" + args.distribution,
            "labels": "int64", "images": imgs, "shape": [imgs, C, H, W],
            "launch": ("cuda_graph replay of the public fwd+bwd call (one graph per input set)" + ("" if world == 1 else " on this rank's images, then one asynchronous all-reduce of the packed partials per step")) if graphs_on else "eager",
            "l2_policy": f"{n_sets} distinct input sets ({n_sets * set_bytes >> 20} MiB > 126 MiB L2) rotated every step",
            "loss": loss_value, "sort_path": args.sort_path,
            "segments_sorted_by_lsd_fallback": fb_segments,
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mpix/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(set_bytes), "d2h_bytes_per_step": 4,
                "h2d_gbs_all_ranks": world * set_bytes / (e2e_ms * 1e-3) / 1e9,
                "note": "per step: pinned host probas+labels -> H2D (side stream, overlapping the previous step's kernels) -> LovaszSoftmax fwd+bwd -> loss D2H; the gradient stays on the device as in training; bound by the host->device copy (h2d_gbs_all_ranks), not by the kernels"},
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic["bytes_per_step"] if traffic else None,
            "traffic_source": traffic["source"] if traffic else None, "algorithmic_bytes": algo_bytes,
            "peak_source": peak_src,
            "what": "whole fwd+bwd step: 160 B/px algorithmic bytes x pixels / CUDA-event step time (the exact sort of 19 x N keys moves ~0.6 KB/px, so the algorithmic fraction is far below each kernel's own DRAM fraction)",
            "design_floor_ms": floor_ms, "fraction_of_floor": floor_ms / ms_step,
            "design_floor_what": f"DRAM floor of the two-movement design: {floor_per_elem:.0f} B per (pixel, class) element + 9 B per pixel at the measured peak",
            "dominant_kernel": dom["name"], "dominant_share_of_step": dom["share_of_step"],
            "dominant_own_gbs": dom["own_gbs"], "dominant_own_frac_of_peak": dom["own_frac_of_peak"],
            "kernels": kernels,
        },
    }
    if parity is not None:
        line["parity_check"] = "ok"
        line["parity_details"] = parity
    if secondary:
        line["secondary"] = secondary

    if not args.no_cpu_baseline and world == 1:
        p_cpu, l_cpu = sets[0][0].cpu(), sets[0][1].cpu()
        if args.hybrid != "none":
            p_cpu = torch.softmax(p_cpu, 1)
        # the same torch ops with device='cuda': the reference's Python as it runs on a GPU (host syncs in
        # unique / nonzero / the empty-class test, ~25 launches per class)
        try:
            pc, lc = p_cpu.to(dev), l_cpu.to(dev)
            for _ in range(2):
                reference_port_step(pc, lc, per_image)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                reference_port_step(pc, lc, per_image)
            torch.cuda.synchronize()
            tc = (time.perf_counter() - t0) / 5
            line["torch_cuda_baseline"] = {"value": px_step / tc / 1e6, "unit": "Mpix/s", "ms_per_step": tc * 1e3, "kind": "port",
                                           "sample": "the whole workload, 2 warm-up + 5 timed steps, wall clock with a final synchronize; "
                                                     "oracle/lovasz_oracle.py (the reference's torch ops, unstable torch.sort) on device='cuda'"}
            del pc, lc
        except Exception as e:  # noqa: BLE001
            line["torch_cuda_baseline"] = {"error": repr(e)[:300]}
        px, times, _ = time_reference_port(p_cpu, l_cpu, per_image, args.cpu_budget_s, 2, 1)
        tb = min(times)
        line["cpu_baseline"] = {
            "value": px / tb / 1e6, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"the whole workload ({imgs}x{C}x{H}x{W}), 1 cold call + {len(times)} timed, best; "
                      "oracle/lovasz_oracle.py torch-CPU restatement of losses/lovasz_loss.py",
        }
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
