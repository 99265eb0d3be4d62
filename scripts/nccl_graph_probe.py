"""Probe: can the sharded per-image step (kernels + one NCCL all-reduce + backward) be captured in a CUDA graph?
Run under torchrun with a timeout; prints eager and graph-replayed ms/step."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtda_semanticsegmentation_b200 import synthetic  # noqa: E402
from rtda_semanticsegmentation_b200.dist import lovasz_softmax_sharded  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
p, l = synthetic.confident_case(4, 512, 1024, seed=3 + rank, device=dev)
pin = p.detach().requires_grad_(True)


def step():
    loss, _ = lovasz_softmax_sharded(pin, l, ignore=255)
    loss.backward()
    return loss


for _ in range(3):
    step(); pin.grad = None
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    step(); pin.grad = None
e1.record(); torch.cuda.synchronize()
print(f"rank {rank} eager {e0.elapsed_time(e1) / 20:.3f} ms/step", flush=True)

side = torch.cuda.Stream(dev)
side.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(side):
    for _ in range(3):
        step(); pin.grad = None
torch.cuda.current_stream(dev).wait_stream(side)
torch.cuda.synchronize(); dist.barrier()
print(f"rank {rank} capturing", flush=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    gl = step()
torch.cuda.synchronize()
print(f"rank {rank} captured", flush=True)
dist.barrier()
for _ in range(3):
    g.replay()
torch.cuda.synchronize(); dist.barrier()
e0.record()
for _ in range(20):
    g.replay()
e1.record(); torch.cuda.synchronize()
print(f"rank {rank} graph {e0.elapsed_time(e1) / 20:.3f} ms/step loss {float(gl):.6f}", flush=True)
dist.destroy_process_group()
