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


def _hybrid_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from rtda_semanticsegmentation_b200 import synthetic
        from rtda_semanticsegmentation_b200.dist import hybrid_ce_lovasz_sharded, shard_range

        b = 2 * world
        labels = synthetic.blocky_labels(b, 64, 96, seed=93, tile=8)
        labels[0, :, :40] = 255                            # ranks see different numbers of valid pixels
        logits = synthetic.confident_logits(labels, seed=93)
        lo, hi = shard_range(b, rank, world)
        res = {}
        for mode, gi in (("global", None), ("rank_mean", b)):
            x = logits[lo:hi].cuda().requires_grad_(True)
            total, ce, lov, stats = hybrid_ce_lovasz_sharded(x, labels[lo:hi].cuda(), ignore=255, lovasz_weight=0.5,
                                                             global_images=gi, ce_reduction=mode)
            (total * 2.0).backward()
            s = torch.stack([total.detach().double(), ce.detach().double(), lov.detach().double()])
            if mode == "rank_mean":
                dist.all_reduce(s)                          # the shares sum to the global values
            res[mode] = (s.cpu().numpy(), x.grad.cpu().numpy(), float(stats["valid_pixels"]))
        # DDP convention: gradients pre-multiplied by the world size
        x = logits[lo:hi].cuda().requires_grad_(True)
        total, _, _, _ = hybrid_ce_lovasz_sharded(x, labels[lo:hi].cuda(), ignore=255, grad_reduction="mean")
        (total * 2.0).backward()
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, f"h{rank}.npz"), lo=lo, hi=hi, g_losses=res["global"][0], g_grad=res["global"][1],
                 g_nv=res["global"][2], r_losses=res["rank_mean"][0], r_grad=res["rank_mean"][1], ddp_grad=x.grad.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_hybrid_ce_lovasz_matches_single_device_composition(tmp_path):
    """BASELINE configs[2] / [4]: CE + 0.5 * per-image Lovasz from logits, batch sharded over the ranks,
    against the reference composition (train_lovasz.py:98-105) on the whole batch."""
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(world, 4)
    mp.spawn(_hybrid_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import lovasz_oracle
    from rtda_semanticsegmentation_b200 import synthetic
    from rtda_semanticsegmentation_b200.dist import shard_range

    b = 2 * world
    labels = synthetic.blocky_labels(b, 64, 96, seed=93, tile=8)
    labels[0, :, :40] = 255
    logits = synthetic.confident_logits(labels, seed=93)

    def composition(lg_full, ce_of):
        lg = lg_full.clone().requires_grad_(True)
        ce = ce_of(lg)
        lv = lovasz_oracle.lovasz_softmax(torch.softmax(lg, 1), labels, ignore=255, per_image=True, stable=True)
        ((ce + 0.5 * lv) * 2.0).backward()
        return float(ce + 0.5 * lv), float(ce), float(lv), lg.grad

    # "global": cross-entropy over all valid pixels of the batch
    gt, gce, glv, ggrad = composition(logits, lambda lg: torch.nn.functional.cross_entropy(lg, labels, ignore_index=255))
    # "rank_mean": mean over ranks of each rank's own cross-entropy mean
    spans = [shard_range(b, r, world) for r in range(world)]
    rt, rce, rlv, rgrad = composition(logits, lambda lg: sum(
        torch.nn.functional.cross_entropy(lg[a:z], labels[a:z], ignore_index=255) for a, z in spans) / world)
    for rank in range(world):
        z = np.load(tmp_path / f"h{rank}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        assert z["g_losses"] == pytest.approx([gt, gce, glv], rel=1e-5)               # every rank holds the global losses
        assert float(z["g_nv"]) == float((labels != 255).sum())
        assert float(np.abs(z["g_grad"] - ggrad[lo:hi].numpy()).max()) <= 1e-5 * float(ggrad.abs().max())
        assert z["r_losses"] == pytest.approx([rt, rce, rlv], rel=1e-5)               # the shares sum to the global values
        assert float(np.abs(z["r_grad"] - rgrad[lo:hi].numpy()).max()) <= 1e-5 * float(rgrad.abs().max())
        assert float(np.abs(z["ddp_grad"] - world * ggrad[lo:hi].numpy()).max()) <= 1e-5 * world * float(ggrad.abs().max())
