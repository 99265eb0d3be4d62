"""CPU, world_size 2 over gloo: the N>1 host logic - per-image sharding with one
all-reduce of the partials, and the confusion-matrix all-reduce.  The per-rank compute
is done by the oracle here (no GPU in this container); on the GPU box the same
functions receive the kernels' outputs (tests/test_multi_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import c_oracle, hist_oracle
from rtda_semanticsegmentation_b200 import synthetic
from rtda_semanticsegmentation_b200.dist import (all_reduce_confusion, losses_from_packed, pack_partials, shard_range,
                                                 sharded_per_image_mean)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        probas, labels = synthetic.confident_case(4, 24, 40, seed=77)
        labels[3] = 255                       # one image with nothing valid
        lo, hi = shard_range(4, rank, world)
        r = c_oracle.lovasz(probas[lo:hi].numpy(), labels[lo:hi].numpy(), 255, per_image=True)
        # per-image losses of the local slice, as the finalize kernel reports them
        cl = torch.from_numpy(r["class_loss"])
        pres = torch.from_numpy(r["present"])
        npres = torch.from_numpy(r["n_present"]).clamp(min=1).float()
        group_loss = torch.stack([cl[i][pres[i] > 0].sum() for i in range(hi - lo)]) / npres
        group_loss.requires_grad_(True)
        loss, stats = sharded_per_image_mean(group_loss, cl, pres)
        loss.backward()
        # fused CE + Lovasz sharded (configs[2], [4]): the buffer k_finalize packs, with this rank's CE sum and
        # valid-pixel count, all-reduced once; the global losses come from the reduced buffer
        logits = synthetic.confident_logits(labels, seed=77)
        lg, lb = logits[lo:hi], labels[lo:hi]
        nv = int((lb != 255).sum())
        ce_sum = float(torch.nn.functional.cross_entropy(lg, lb, ignore_index=255, reduction="sum").double()) if nv else 0.0
        packed = pack_partials(group_loss.detach(), cl, pres, ce_sum=ce_sum, n_valid=nv)
        dist.all_reduce(packed)
        h_lov, h_ce, h_nv = losses_from_packed(packed, 19)
        lab, prd = synthetic.validation_maps(6, 32, 32, seed=78)
        lo2, hi2 = shard_range(6, rank, world)
        h = torch.from_numpy(hist_oracle.validation_confusion(lab[lo2:hi2].numpy(), prd[lo2:hi2].numpy(), 19, 255))
        all_reduce_confusion(h)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), loss=loss.detach().numpy(), grad=group_loss.grad.numpy(),
                 hist=h.numpy(), images=stats["images"].numpy(), counts=stats["class_count"].numpy(),
                 h_lov=h_lov.numpy(), h_ce=h_ce.numpy(), h_nv=h_nv.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_sharding(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    probas, labels = synthetic.confident_case(4, 24, 40, seed=77)
    labels[3] = 255
    want = c_oracle.lovasz(probas.numpy(), labels.numpy(), 255, per_image=True)
    lab, prd = synthetic.validation_maps(6, 32, 32, seed=78)
    want_hist = hist_oracle.validation_confusion(lab.numpy(), prd.numpy(), 19, 255)
    logits = synthetic.confident_logits(labels, seed=77)
    want_ce = float(torch.nn.functional.cross_entropy(logits, labels, ignore_index=255))   # train_lovasz.py:98 on the whole batch
    for rank in range(world):
        z = np.load(tmp_path / f"r{rank}.npz")
        assert float(z["h_ce"]) == pytest.approx(want_ce, rel=1e-6)                  # CE over ALL ranks' valid pixels
        assert float(z["h_lov"]) == pytest.approx(float(want["loss"]), rel=1e-6)
        assert float(z["h_nv"]) == float((labels != 255).sum())
        assert float(z["loss"]) == pytest.approx(float(want["loss"]), rel=1e-6)      # every rank holds the global loss
        assert np.allclose(z["grad"], 1.0 / 4.0)                                     # d loss / d loss_b = 1 / B_global
        assert np.array_equal(z["hist"], want_hist)                                  # bit-exact counts
        assert float(z["images"]) == 4.0
        assert np.array_equal(z["counts"], want["present"].sum(0).astype(np.float64))
