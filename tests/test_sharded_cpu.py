"""Host-side logic of the batch-sharded loss wrapper on CPU: world_size-2 `gloo` ranks against the
single-process result.  The three CUDA entry points of EarthMoverDistance are replaced by the CPU oracle
in a test-only subclass; everything else (histogram/ratio all-reduces, value/gradient contract, shard
bounds, `.log` forwarding) is the product code."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from pointcloud_b200 import synth  # noqa: E402
from pointcloud_b200.losses import (ChamferDistance, EarthMoverDistance, FilterClasses, FilteringChamferDistance,  # noqa: E402
                                    SegmentingChamferDistance)
from pointcloud_b200.sharded import ShardedLoss, shard_bounds  # noqa: E402


class CpuEMD(EarthMoverDistance):
    """EarthMoverDistance with its kernel entry points served by the CPU oracle / plain torch (tests only)."""

    def _auction(self, pred, target, want_epilogue=False):
        xyz1, xyz2 = pred[:, :, :3], target[:, :, :3]
        r = oracle.emd_forward(xyz1, xyz2, self.eps, self.iterations)
        return xyz1, xyz2, torch.from_numpy(r["dist"]), torch.from_numpy(r["assignment"])

    def _matched_hist(self, target, assignment):
        lab = target[:, :, 3].long().take_along_dim(assignment.long(), 1)
        return torch.bincount(lab.view(-1), minlength=self.C), lab.int()

    def _point_sums(self, xyz1, xyz2, dists, assignment, matched, class_weights):
        m = xyz2.take_along_dim(assignment.long().unsqueeze(-1), 1)
        d = ((xyz1 - m) ** 2).sum(-1)
        w = torch.ones_like(d) if class_weights is None else class_weights[matched.long()]
        return torch.stack([(d.sqrt() * w).sum(), w.sum()])


    def _ce_sums(self, pred, matched, class_weights):
        logp = torch.log_softmax(pred[:, :, 3:].float(), dim=2)
        nll = -logp.gather(2, matched.long().unsqueeze(-1)).squeeze(-1)
        w = class_weights[matched.long()]
        pred_hist = torch.bincount(pred[:, :, 3:].argmax(dim=2).view(-1), minlength=self.C)
        return torch.stack([(nll * w).sum(), w.sum()]), pred_hist

    def _mse_sums(self, pred, target, assignment):
        diff = pred[:, :, 3:] - target[:, :, 3:].take_along_dim(assignment.long().unsqueeze(-1), 1)
        return torch.stack([(diff * diff).sum(), torch.tensor(float(diff.numel()))])


class CpuChamfer(ChamferDistance):
    def __call__(self, pred, target):
        from oracle import loss_oracle
        return loss_oracle.chamfer_distance(pred, target)[0]


class CpuFilteringChamfer(FilteringChamferDistance):
    """Product filter + pad logic (torch path on a CPU tensor), Chamfer arithmetic from the CPU oracle."""

    def __call__(self, pred, target):
        from oracle import loss_oracle
        tgt, num = self._filter_pad(target, torch.float32)
        return loss_oracle.chamfer_distance(pred.to(torch.float32), tgt, y_lengths=num)[0]


class CpuSegmentingChamfer(SegmentingChamferDistance):
    def __init__(self, class_labels):
        self.classs_losses = {c: CpuFilteringChamfer(FilterClasses([l], label_dim=3)) for c, l in class_labels.items()}


def _data(kind):
    if kind == "seg":
        return synth.segmenter_batch(4, 256, seed=21, regime="noisy")
    if kind == "ae":
        return synth.autoencoder_batch(4, 256, seed=22)
    x, y = synth.uniform_clouds(4, 200, seed=23)
    return x, y


def _make(kind):
    if kind == "seg":
        return CpuEMD(0.005, 50, num_classes=5)
    if kind == "ae":
        return CpuEMD(0.005, 50)
    return CpuChamfer()


def _worker(rank, world, port, kind, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pred, target = _data(kind)
        lo, hi = shard_bounds(pred.shape[0], world, rank)
        p = pred[lo:hi].clone().requires_grad_()
        fn = ShardedLoss(_make(kind))
        logged = {}
        fn.log = lambda k, v: logged.__setitem__(k, float(v.detach()))
        loss = fn(p, target[lo:hi])
        loss.backward()
        q.put((rank, float(loss.detach()), p.grad.numpy(), logged))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["ae", "seg", "chamfer"])
def test_sharded_loss_two_gloo_ranks_match_single_process(kind):
    world = 2
    pred, target = _data(kind)
    p = pred.clone().requires_grad_()
    single = _make(kind)
    ref = single(p, target)
    ref.backward()

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, loss, grad, logged in got:
        assert loss == pytest.approx(float(ref.detach()), rel=2e-6)                     # GLOBAL value on every rank
        lo, hi = shard_bounds(pred.shape[0], world, rank)
        # backward = world * d(global)/d(local): DDP's gradient averaging then yields the single-GPU step
        np.testing.assert_allclose(grad / world, p.grad[lo:hi].numpy(), rtol=3e-5, atol=1e-10)
        if kind != "chamfer":
            assert "train_loss/EMD" in logged and "train_loss/feature" in logged       # .log forwarded (train.py:161)


def test_shard_bounds_cover_batch_contiguously():
    for b in (1, 4, 25, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(b, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_sharded_loss_without_process_group_is_the_plain_loss():
    pred, target = _data("seg")
    p1, p2 = pred.clone().requires_grad_(), pred.clone().requires_grad_()
    a = CpuEMD(0.005, 50, num_classes=5)(p1, target)
    b = ShardedLoss(CpuEMD(0.005, 50, num_classes=5))(p2, target)
    a.backward(); b.backward()
    assert float(a.detach()) == pytest.approx(float(b.detach()), rel=1e-6)
    np.testing.assert_allclose(p1.grad.numpy(), p2.grad.numpy(), rtol=1e-5, atol=1e-10)
    with pytest.raises(TypeError):
        ShardedLoss(lambda p, t: p.sum())
