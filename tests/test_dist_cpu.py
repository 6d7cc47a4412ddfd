"""The N > 1 path on CPU: world_size 2, gloo backend, the CPU stand-in ops (tests/fake_ops.py).

SURVEY.md 8e: each rank trains on its own labeled + unlabeled mini-batch; the only exchange is ONE all-reduce (sum) of
the student's flat gradient, scaled by 1 / world inside the optimizer kernel; the teacher is the EMA of bit-identical
student weights, so it needs no collective.  Checked here: (1) after a step all ranks hold identical student and teacher
weights, (2) the step equals the mean of the two single-rank steps (first step, zero momentum: the update is linear in
the gradient), (3) BatchNorm running statistics stay rank-local."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Patch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir, buckets):
    os.environ["B200_DP_BUCKETS"] = buckets          # "1": decoder / encoder gradient buckets (trainers.py), "0": one all-reduce
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import fake_ops
    fake_ops.install(_Patch())
    from cv_ssl_mis_b200.networks import unet as unet_mod
    from cv_ssl_mis_b200.trainers import MeanTeacherTrainer

    B, Lb, H, W, it0 = 4, 2, 32, 32, 1200

    def build():
        torch.manual_seed(123)                       # identical initial weights on every rank (the product broadcasts rank 0's)
        return unet_mod.UNet(1, 4, seed=11), unet_mod.UNet(1, 4, seed=22)

    def batch(r):
        g = torch.Generator().manual_seed(1000 + r)
        return torch.rand(B, 1, H, W, generator=g), torch.randint(0, 4, (B, H, W), generator=g).to(torch.uint8)

    def run(r, pg):
        s, t = build()
        tr = MeanTeacherTrainer(s, t, batch_size=B, labeled_bs=Lb, patch_size=(H, W), start_iter=it0, noise_seed=50 + r,
                                process_group=pg)
        # per-rank dropout epochs differ in the product (seed + rank); here the same seeds keep the check linear
        tr.step(*batch(r))
        return tr, s, t

    solo = [run(r, None) for r in range(world)]      # both single-rank steps, computed locally on every rank
    tr, s, t = run(rank, dist.group.WORLD)
    flat, ema = tr.flat.data.clone(), tr.ema_flat.data.clone()
    expect = sum(x[0].flat.data for x in solo) / world
    torch.testing.assert_close(flat, expect, rtol=1e-5, atol=1e-7)
    # EMA at it0 = 1200: alpha = 0.99 of the teacher's own init + 0.01 of the new student
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(gathered[0], g) for g in gathered), "student weights diverged across ranks"
    gathered_e = [torch.empty_like(ema) for _ in range(world)]
    dist.all_gather(gathered_e, ema)
    assert all(torch.equal(gathered_e[0], g) for g in gathered_e), "teacher weights diverged across ranks"
    # BatchNorm running stats are rank-local (different batches): they must differ between the ranks
    rm = s.encoder.in_conv.conv_conv[1].running_mean.clone()
    both = [torch.empty_like(rm) for _ in range(world)]
    dist.all_gather(both, rm)
    assert not torch.equal(both[0], both[1])
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("buckets", ["0", "1"])
def test_two_rank_gradient_allreduce_gloo(tmp_path, buckets):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), buckets), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def _ct_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import fake_ops
    fake_ops.install(_Patch())
    from cv_ssl_mis_b200.networks import unet as unet_mod
    from cv_ssl_mis_b200.trainers import CrossTeachingTrainer

    B, Lb, P, it0 = 4, 2, 32, 3000

    def run(r, pg):
        torch.manual_seed(321)
        m1, m2 = unet_mod.UNet(1, 4, seed=11), unet_mod.UNet(1, 4, seed=22)
        tr = CrossTeachingTrainer(m1, m2, batch_size=B, labeled_bs=Lb, patch_size=(P, P), num_classes=4, start_iter=it0,
                                  pseudo_loss="ce", process_group=pg)
        g = torch.Generator().manual_seed(2000 + r)
        tr.step(torch.rand(B, 1, P, P, generator=g), torch.randint(0, 4, (B, P, P), generator=g).to(torch.uint8))
        return tr

    solo = [run(r, None) for r in range(world)]
    tr = run(rank, dist.group.WORLD)
    for i in range(2):                                 # one all-reduce per model: both stay replicated, update = mean of solos
        flat = tr.flats[i].data
        torch.testing.assert_close(flat, sum(s.flats[i].data for s in solo) / world, rtol=1e-5, atol=1e-7)
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat.clone())
        assert torch.equal(both[0], both[1])
    open(os.path.join(out_dir, f"ct_ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_cross_pseudo_supervision_gloo(tmp_path):
    world = 2
    mp.spawn(_ct_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ct_ok{r}").exists() for r in range(world))
