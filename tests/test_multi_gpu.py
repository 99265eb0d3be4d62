"""GPU, >= 2 devices: the sharded paths over NCCL - per-image Lovasz with ONE all-reduce of the
partials, confusion matrix with one all-reduce of the counts - against the single-GPU result and
the oracle.  Skipped on a 1-GPU box; the packing logic is covered on CPU by tests/test_dist_gloo.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import rtda_semanticsegmentation_b200 as rtda
        from rtda_semanticsegmentation_b200 import synthetic
        from rtda_semanticsegmentation_b200.dist import lovasz_softmax_sharded, shard_range

        b = 2 * world
        probas, labels = synthetic.confident_case(b, 96, 128, seed=91)
        labels[b - 1] = 255                                  # one image with nothing valid
        lo, hi = shard_range(b, rank, world)
        p = probas[lo:hi].cuda().requires_grad_(True)
        loss, stats = lovasz_softmax_sharded(p, labels[lo:hi].cuda(), ignore=255)
        (loss * 3.0).backward()
        # known global batch size: asynchronous collective, the forward returns this rank's share
        p2 = probas[lo:hi].cuda().requires_grad_(True)
        share, stats2 = lovasz_softmax_sharded(p2, labels[lo:hi].cuda(), ignore=255, global_images=b)
        (share * 3.0).backward()
        share_sum = share.detach().double().clone()
        dist.all_reduce(share_sum)
        lab, prd = synthetic.validation_maps(3 * world, 128, 160, seed=92)
        lo2, hi2 = shard_range(3 * world, rank, world)
        cm = rtda.ConfusionMatrix(19, 255)
        for i in range(lo2, hi2):
            cm.update(lab[i].cuda(), prd[i].cuda())
        cm.all_reduce()
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), loss=loss.detach().cpu().numpy(), grad=p.grad.cpu().numpy(),
                 hist=cm.compute(), images=stats["images"].cpu().numpy(), lo=lo, hi=hi,
                 grad2=p2.grad.cpu().numpy(), share_sum=share_sum.cpu().numpy(), loss2=stats2["loss"].cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_paths_match_single_gpu_and_oracle(tmp_path):
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(world, 4)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import c_oracle, hist_oracle
    from rtda_semanticsegmentation_b200 import synthetic

    b = 2 * world
    probas, labels = synthetic.confident_case(b, 96, 128, seed=91)
    labels[b - 1] = 255
    want = c_oracle.lovasz(probas.numpy(), labels.numpy(), 255, per_image=True, grad_out=3.0)
    lab, prd = synthetic.validation_maps(3 * world, 128, 160, seed=92)
    want_hist = hist_oracle.validation_confusion(lab.numpy(), prd.numpy(), 19, 255)
    gmax = np.abs(want["grad"]).max()
    for rank in range(world):
        z = np.load(tmp_path / f"r{rank}.npz")
        assert float(z["loss"]) == pytest.approx(float(want["loss"]), rel=2e-6)       # every rank has the global loss
        assert float(z["images"]) == b
        lo, hi = int(z["lo"]), int(z["hi"])
        # the local scale is (3/B_global)/K computed in a different association order than the oracle's: 1-2 ulp
        assert np.abs(z["grad"] - want["grad"][lo:hi]).max() <= 1e-6 * gmax
        assert np.array_equal(z["hist"], want_hist)                                   # counts are exact
        # asynchronous mode: the shares sum to the global loss, the lazily reduced loss equals it, same gradients
        assert float(z["share_sum"]) == pytest.approx(float(want["loss"]), rel=2e-6)
        assert float(z["loss2"]) == pytest.approx(float(want["loss"]), rel=2e-6)
        assert np.abs(z["grad2"] - want["grad"][lo:hi]).max() <= 1e-6 * gmax
