#!/usr/bin/env python
"""Benchmark of the Lovasz-Softmax fwd+bwd hot path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4|c5]

One "step" = one forward + backward of the loss over one batch of synthetic input
(`LovaszSoftmax(ignore=255)(probas, labels)`; `loss.backward()`).  Prints ONE JSON line.

* N = 1 (default): workload c2 = BASELINE.json configs[1], 4x19x512x1024, global
  (per_image=False), classes='present', D_conf synthetic distribution.
* N > 1 (torchrun, one rank per GPU): the path shards per image, so every rank runs
  per_image=True on its own 4x19x512x1024 slice of a 4N-image batch and ONE NCCL
  all-reduce of the 40-double partial buffer gives the global loss (weak scaling).
* `value`   - Mpix/s (pixels = B*H*W incl. ignored) with inputs resident in HBM; several
  input sets larger than L2 are rotated so no step finds its inputs in L2.
* `e2e`     - the same through the public API from pinned HOST buffers: H2D of probas and
  labels and D2H of the loss inside the timed region.
* `roofline` - algorithmic bytes (160 B/px) / step time against the measured
  HBM peak of MEASURED_PEAKS.json; `kernels` lists every kernel's live CUDA-event time.
* `cpu_baseline` - the oracle port of the reference's CPU path timed on this box's cores.
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

ALGO_BYTES_PER_PX = 160.0   # read 19 fp32 probs + 1 int64 label, write 19 fp32 grads
CM_BYTES_PER_PX = 16.0      # int64 label + int64 prediction

WORKLOADS = {
    # name: (images per GPU, H, W, per_image at N=1, description)
    "c1": (1, 512, 1024, False, "configs[0]: 1x19x512x1024 global"),
    "c2": (4, 512, 1024, False, "configs[1]: 4x19x512x1024 global, classes=present"),
    "c3": (16, 720, 1280, True, "configs[2]: 16x19x720x1280 per_image (Lovasz only), batch sharded over ranks"),
    "c5": (8, 512, 1024, False, "configs[4]: 8x19x512x1024 global"),
    # configs[3] is the confusion-matrix path (its own metric): see bench_confusion()
    "c4": (1, 1024, 2048, False, "configs[3]: validation confusion matrix, 1024x2048 int64 label + prediction maps, images sharded over ranks"),
}


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


def reference_port_step(probas_cpu, labels_cpu, per_image):
    """One fwd+bwd of the reference's CPU path as restated op for op in oracle/lovasz_oracle.py
    (unstable torch.sort exactly like lovasz_loss.py:173)."""
    from oracle import lovasz_oracle

    lovasz_oracle.lovasz_softmax_with_grad(probas_cpu, labels_cpu, ignore=255, per_image=per_image, stable=False)


def time_reference_port(probas_cpu, labels_cpu, per_image, budget_s, steps, warmup):
    """Times the oracle port on a bounded sample: as many leading images as fit the budget."""
    torch.set_num_threads(os.cpu_count() or 1)
    b = probas_cpu.size(0)
    t0 = time.perf_counter()
    reference_port_step(probas_cpu[:1], labels_cpu[:1], per_image)
    t1 = time.perf_counter() - t0
    imgs = max(1, min(b, int(budget_s / max(t1, 1e-3) / max(steps + warmup, 1))))
    p, l = probas_cpu[:imgs].contiguous(), labels_cpu[:imgs].contiguous()
    for _ in range(warmup):
        reference_port_step(p, l, per_image)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        reference_port_step(p, l, per_image)
        times.append(time.perf_counter() - t0)
    px = imgs * p.size(2) * p.size(3)
    return px, times, imgs


def bench_confusion(args, rank, local_rank, world):
    """configs[3]: one step = one 1024x2048 validation map pair accumulated into the 19x19 matrix
    (validation.py:98-109).  Maps are sharded over ranks; ONE all-reduce of the counts at the end,
    inside the timed region.  16 algorithmic bytes per pixel."""
    import numpy as np

    import rtda_semanticsegmentation_b200 as rtda
    from rtda_semanticsegmentation_b200 import _lib, synthetic

    _, H, W, _, desc = WORKLOADS["c4"]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n_sets = 10                                             # 10 x 33.5 MB = 335 MB > L2
    maps = [synthetic.validation_maps(1, H, W, seed=1234 + 400 + rank + 100 * i, device=dev) for i in range(n_sets)]
    px = H * W
    cm = rtda.ConfusionMatrix(19, 255, device=dev)
    warmup = max(args.warmup, 3)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        cm.update(*maps[i % n_sets])
    cm.reset()
    # one CUDA graph holding the n_sets per-map update calls (a validation loop's launch pattern
    # without the ~19 us of Python per call); eager launches if capture is not possible
    graph = None
    if world == 1 and not args.no_graph:
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
    steps = args.steps if graph is None else max(n_sets, args.steps // n_sets * n_sets)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if graph is not None:
        for i in range(steps // n_sets):
            graph.replay()
    else:
        for i in range(steps):
            cm.update(*maps[i % n_sets])
    cm.all_reduce()
    e1.record()
    sync()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    value = world * px / (ms_step * 1e-3) / 1e6
    total = int(cm.compute().sum())
    # per-launch kernel time from the library's CUDA events, eager launches
    _lib.profile_enable(True)
    _lib.profile_read()
    for i in range(args.steps):
        cm.update(*maps[i % n_sets])
    stages = _lib.profile_read()
    _lib.profile_enable(False)

    # end to end through the reference signature: numpy maps on the host -> fast_hist -> numpy matrix
    host = [(l.cpu().numpy().ravel(), p.cpu().numpy().ravel()) for l, p in maps[:3]]
    for i in range(3):
        rtda.fast_hist(*host[i % 3], 19)
    sync()
    t0 = time.perf_counter()
    acc = np.zeros((19, 19))
    for i in range(args.steps):
        acc += rtda.fast_hist(*host[i % 3], 19)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peak, peak_src = peaks()
    k_ms, k_n = stages.get("k_confusion", (0.0, 1))
    per = k_ms / max(k_n, 1)
    achieved = CM_BYTES_PER_PX * px / (per * 1e-3) / 1e9
    line = {
        "metric": "confusion_matrix_mpix_per_s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": f"c4: {desc}", "classes": 19, "ignore": 255, "maps_per_step": 1,
                   "launch": "cuda_graph replay of 10 per-map update calls" if graph is not None else "eager",
                   "l2_policy": f"{n_sets} distinct map pairs ({n_sets * px * 16 >> 20} MiB > 126 MiB L2) rotated every step",
                   "counted_pixels": total},
        "clocks": clocks,
        "e2e": {"value": world * px / (e2e_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": px * 16, "d2h_bytes_per_step": 19 * 19 * 8,
                "note": "rtda.fast_hist(label_true_np, label_pred_np, 19): pageable numpy maps -> H2D -> kernel -> 19x19 int64 D2H, wall clock"},
        "gpu_launches": int(k_n),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "dominant_kernel": "k_confusion",
                     "ms_per_launch": per, "what": "16 B/px (int64 label + int64 prediction) x 2097152 px / k_confusion CUDA-event time"},
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle import hist_oracle
        l, p = maps[0][0].cpu().numpy(), maps[0][1].cpu().numpy()
        hist_oracle.validation_confusion(l, p, 19, 255)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            hist_oracle.validation_confusion(l, p, 19, 255)
            ts.append(time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": px / min(ts) / 1e6, "unit": "Mpix/s", "cores": 1, "kind": "port",
                                "sample": "one 1024x2048 map, best of 5; oracle/hist_oracle.py numpy restatement of "
                                          "validation.py:98-109 + utils.py:62-89 (np.bincount is single-threaded)"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--distribution", default="conf", choices=["conf", "diffuse"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--hybrid", default="none", choices=["none", "fused", "unfused"],
                    help="time CE + 0.5*Lovasz from LOGITS (train_lovasz.py:98-108): fused = rtda_hybrid_fwd/bwd, "
                         "unfused = torch softmax + torch CE + LovaszSoftmax; not the BASELINE metric")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    ap.add_argument("--sort-path", default="bucket", choices=["bucket", "lsd"],
                    help="forward sort path (A/B aid): sample-partitioned bucket sort (default) or LSD radix sort everywhere")
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "c4" and args.impl == "ours":
        return bench_confusion(args, rank, local_rank, world)
    imgs, H, W, per_image_n1, desc = WORKLOADS[args.workload]
    C = 19
    from rtda_semanticsegmentation_b200 import synthetic

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        gen = synthetic.confident_case if args.distribution == "conf" else synthetic.diffuse_case
        if args.workload == "c4":
            import numpy as np
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
        cfg_id = {"c1": 1, "c2": 2, "c3": 3, "c5": 5}[args.workload]
        probas, labels = gen(min(imgs, 4), H, W, seed=1234 + 100 * cfg_id)
        per_image = per_image_n1 or world > 1
        px, times, used = time_reference_port(probas, labels, per_image, 150.0, args.steps, args.warmup)
        t = sum(times) / len(times)
        val = px / t / 1e6
        line = {
            "impl": "reference", "metric": "lovasz_softmax_fwd_bwd_mpix_per_s", "value": val, "unit": "Mpix/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "classes": C, "distribution": "D_" + args.distribution,
                       "per_image": per_image},
            "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"first {used} image(s) of the workload per step ({used}x{C}x{H}x{W}), "
                                       "oracle/lovasz_oracle.py torch-CPU restatement of losses/lovasz_loss.py"},
            "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import rtda_semanticsegmentation_b200 as rtda
    from rtda_semanticsegmentation_b200 import _lib
    from rtda_semanticsegmentation_b200.dist import lovasz_softmax_sharded

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    rtda.set_sort_path(args.sort_path)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    per_image = per_image_n1 or world > 1
    if args.workload == "c3":
        imgs = max(1, imgs // world)       # configs[2]: strong split of the 16-image batch
        scaling = "strong"
    else:
        scaling = "weak"
    cfg_id = {"c1": 1, "c2": 2, "c3": 3, "c5": 5}[args.workload]
    gen = synthetic.confident_case if args.distribution == "conf" else synthetic.diffuse_case

    # several distinct input sets, together larger than L2 (126 MB), rotated every step
    set_bytes = imgs * H * W * (4 * C + 8)
    n_sets = max(2, min(6, -(-400 * 1024 * 1024 // set_bytes)))
    sets = []
    for i in range(n_sets):
        if args.hybrid != "none":       # the hybrid step starts from logits
            l = synthetic.blocky_labels(imgs, H, W, seed=1234 + 100 * cfg_id + rank + 1000 * i, device=dev)
            p = synthetic.confident_logits(l, seed=1234 + 100 * cfg_id + rank + 1000 * i)
        else:
            p, l = gen(imgs, H, W, seed=1234 + 100 * cfg_id + rank + 1000 * i, device=dev)
        sets.append((p.contiguous(), l.contiguous()))
    px_step = imgs * H * W
    lovasz = rtda.LovaszSoftmax(ignore=255, per_image=per_image)
    fused = rtda.HybridCELovasz(ignore=255, lovasz_weight=0.5, per_image=per_image)
    if args.hybrid != "none" and world > 1:
        raise SystemExit("--hybrid is a single-GPU measurement")

    def crit(x, l):
        if args.hybrid == "fused":
            return fused(x, l)
        if args.hybrid == "unfused":
            return torch.nn.functional.cross_entropy(x, l, ignore_index=255) + 0.5 * lovasz(torch.softmax(x, dim=1), l)
        return lovasz(x, l)

    def step(p, l):
        p = p.detach().requires_grad_(True)
        if world > 1:
            # fixed per-rank batch: the global image count is known, the all-reduce is asynchronous and
            # the step's loss is this rank's share of the global mean (dist.lovasz_softmax_sharded)
            loss, _ = lovasz_softmax_sharded(p, l, ignore=255, global_images=imgs * world)
        else:
            loss = crit(p, l)
        loss.backward()
        return loss, p.grad

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(*sets[i % n_sets])
    barrier()

    # ---- CUDA graphs of one fwd+bwd per input set (single GPU): the launch path of the timed region
    # is then free of host jitter.  Falls back to eager launches if capture is not possible.
    graphs = None
    # N > 1: capturing the NCCL all-reduce itself hung (scripts/nccl_graph_probe.py), so the graph holds
    # the LOCAL part of the sharded step - per-image loss of this rank's images times B_local/B_global,
    # which is this rank's share of the global mean and has exactly the sharded gradient - and the one
    # all-reduce of the packed partials is issued asynchronously after every replay (the same
    # composition dist.lovasz_softmax_sharded(global_images=...) performs).
    share = 1.0 / world
    pending = [None]

    def reduce_partials():
        if pending[0] is not None:
            pending[0].wait()                      # stream-side wait; at most one collective in flight
        pending[0] = dist.all_reduce(lovasz.last_outputs.packed, op=dist.ReduceOp.SUM, async_op=True)

    if not args.no_graph:
        try:
            graphs = []
            side = torch.cuda.Stream(dev)
            for p, l in sets:
                pin = p.detach().requires_grad_(True)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(2):
                        (crit(pin, l) * share).backward()
                        pin.grad = None
                torch.cuda.current_stream(dev).wait_stream(side)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    gl = crit(pin, l) * share if world > 1 else crit(pin, l)
                    gl.backward()
                graphs.append((g, gl, pin, lovasz.last_outputs))
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"[bench] CUDA graph capture failed, timing eager launches: {e!r}", file=sys.stderr)
            graphs = None

    # ---- timed region 1: device-resident inputs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    host_t0 = time.perf_counter()
    if graphs is not None:
        for i in range(args.steps):
            graphs[i % n_sets][0].replay()
            if world > 1:
                lovasz.last_outputs = graphs[i % n_sets][3]
                reduce_partials()
        loss = graphs[(args.steps - 1) % n_sets][1]
        if pending[0] is not None:
            pending[0].wait()
    else:
        for i in range(args.steps):
            loss, grad = step(*sets[i % n_sets])
    host_ms_step = (time.perf_counter() - host_t0) * 1e3 / args.steps   # host enqueue time, no sync inside
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * px_step / (ms_step * 1e-3) / 1e6
    loss_value = float(loss.item())

    # ---- timed region 1b: the same steps launched eagerly with the library's per-launch CUDA events on
    # (kernel table of the roofline; also the eager step time)
    _lib.profile_enable(True)
    _lib.profile_read()
    barrier()
    e0.record()
    for i in range(args.steps):
        step(*sets[i % n_sets])
    e1.record()
    barrier()
    eager_ms_step = e0.elapsed_time(e1) / args.steps
    stages = _lib.profile_read()
    _lib.profile_enable(False)

    # ---- timed region 2: end to end from pinned host buffers.  Every step's probas + labels are
    # copied host -> device inside the region (double-buffered on a side stream so the copy of step
    # i+1 overlaps the kernels of step i) and every step's loss is copied back to the host.
    from rtda_semanticsegmentation_b200.host_pipeline import DeviceStager

    hp = [(p.cpu().pin_memory(), l.cpu().pin_memory()) for p, l in sets[:2]]
    stager = DeviceStager(sets[0], dev, depth=2)
    host_loss = torch.empty((args.steps + 8,), dtype=torch.float32).pin_memory()

    def e2e_run(n):
        slot = stager.submit(hp[0])
        for i in range(n):
            nxt = stager.submit(hp[(i + 1) % 2]) if i + 1 < n else None
            dp, dl = stager.acquire(slot)
            loss, _ = step(dp, dl)
            stager.release(slot)
            host_loss[i].copy_(loss.detach(), non_blocking=True)
            slot = nxt

    e2e_run(3)
    barrier()
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2.item()) / args.steps
    e2e_value = world * px_step / (e2e_ms * 1e-3) / 1e6
    assert abs(float(host_loss[args.steps - 1]) - float(step(*sets[(args.steps - 1) % 2])[0])) < 1e-6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
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
    own_bytes = {
        "memset_workspace": None,
        "k_prep": px_step * (8 + 1),                       # labels in, class codes out
        "k_split": E // 16 * 8,                            # 1/16 sample, half of every 32-byte sector used
        "k_part": E * 4 + E * 1 + E * 8,                   # probabilities + codes (re-read per class) in, pairs out
        "k_plan": None,
        "k_local": E * 8 + E * 4,                          # pairs in, 4-byte gradient scatter out
        "k_hist_fb": None,
        "k_scan": None,
        "k_sort_pass<0,probas>": E * 4 + px_step * 1 + E * 8,
        "k_sort_pass<1>": E * 16,
        "k_sort_pass<2>": E * 16,
        "k_rank_last": E * 8 + E * 4,
        "k_finalize": None,
        "k_bwd": E * 8,
    }
    fb_segments = int(rtda.segment_sort_paths((imgs, C, H, W), per_image).sum())
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
                        "share_of_step": round(ms / args.steps / eager_ms_step, 4),
                        "own_gbs": round(ob / (per * 1e-3) / 1e9, 1) if ob and per > 0 else None,
                        "own_frac_of_peak": round(ob / (per * 1e-3) / 1e9 / peak, 4) if ob and per > 0 else None})
    kernels.sort(key=lambda k: -k["share_of_step"])
    dom = kernels[0]
    gaps = stages.get("gaps_between_launches", (0.0, 0))

    line = {
        "metric": "lovasz_softmax_fwd_bwd_mpix_per_s" if args.hybrid == "none" else f"hybrid_ce_lovasz_fwd_bwd_mpix_per_s[{args.hybrid}]",
        "value": value, "unit": "Mpix/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step, "host_enqueue_ms_per_step": host_ms_step,
        "eager_ms_per_step": eager_ms_step, "eager_gap_ms_per_step": gaps[0] / args.steps, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"{args.workload}: {desc}" + ("" if world == 1 else f"; per rank {imgs}x{C}x{H}x{W}, per_image=True sharded over {world} ranks, 1 all-reduce of 40 doubles"),
            "classes": C, "ignore": 255, "per_image": per_image, "distribution": "D_" + args.distribution,
            "labels": "int64", "launch": ("cuda_graph replay of the public fwd+bwd call (one graph per input set)" + ("" if world == 1 else " on this rank's images, then one asynchronous all-reduce of the packed partials per step")) if graphs is not None else "eager",
            "l2_policy": f"{n_sets} distinct input sets ({n_sets * set_bytes >> 20} MiB > 126 MiB L2) rotated every step",
            "loss": loss_value, "sort_path": args.sort_path,
            "segments_sorted_by_lsd_fallback": fb_segments,
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mpix/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(set_bytes), "d2h_bytes_per_step": 4,
                "note": "per step: pinned host probas+labels -> H2D (side stream, overlapping the previous step's kernels) -> LovaszSoftmax fwd+bwd -> loss D2H; the gradient stays on the device as in training"},
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic["bytes_per_step"] if traffic else None,
            "traffic_source": traffic["source"] if traffic else None, "algorithmic_bytes": algo_bytes,
            "peak_source": peak_src,
            "what": "whole fwd+bwd step: 160 B/px algorithmic bytes x pixels / CUDA-event step time (the exact sort of 19 x N keys moves ~0.6 KB/px, so the algorithmic fraction is far below each kernel's own DRAM fraction)",
            "dominant_kernel": dom["name"], "dominant_share_of_step": dom["share_of_step"],
            "dominant_own_gbs": dom["own_gbs"], "dominant_own_frac_of_peak": dom["own_frac_of_peak"],
            "kernels": kernels,
        },
    }

    if not args.no_cpu_baseline and world == 1:
        p_cpu, l_cpu = sets[0][0].cpu(), sets[0][1].cpu()
        px, times, used = time_reference_port(p_cpu, l_cpu, per_image, args.cpu_budget_s, 2, 1)
        tb = min(times)
        line["cpu_baseline"] = {
            "value": px / tb / 1e6, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"first {used} image(s) of the workload ({used}x{C}x{H}x{W}), 1 warm-up + 2 timed, best; "
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
