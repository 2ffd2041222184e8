"""Batch-sharded loss on real GPUs over NCCL (needs >= 2 devices; skipped on a 1-GPU box): every rank evaluates its
shard with the CUDA kernels, the wrapper all-reduces the batch statistics, and value + DDP-scaled gradients must
equal the single-GPU loss on the whole batch."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, kind, out_dir):
    import pointcloud_b200 as pcl
    from pointcloud_b200 import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        pred, target, fn = _case(pcl, synth, kind)
        lo, hi = pcl.shard_bounds(pred.shape[0], world, rank)
        p = pred[lo:hi].cuda().requires_grad_()
        loss = pcl.ShardedLoss(fn)(p, target[lo:hi].cuda())
        loss.backward()
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), loss=float(loss.detach()), grad=p.grad.cpu().numpy())
    finally:
        dist.destroy_process_group()


def _case(pcl, synth, kind):
    if kind == "seg":
        pred, target = synth.segmenter_batch(8, 1024, seed=31, regime="noisy")
        return pred, target, pcl.EarthMoverDistance(0.005, 50, num_classes=5)
    if kind == "ae":
        pred, target = synth.autoencoder_batch(8, 1024, seed=32)
        return pred, target, pcl.EarthMoverDistance(0.005, 50)
    x, y = synth.uniform_clouds(8, 777, seed=33)
    return x, y, pcl.ChamferDistance()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("kind", ["ae", "seg", "chamfer"])
def test_sharded_loss_nccl_two_gpus(kind, tmp_path):
    import pointcloud_b200 as pcl
    from pointcloud_b200 import synth
    world = 2
    pred, target, fn = _case(pcl, synth, kind)
    p = pred.cuda().requires_grad_()
    ref = fn(p, target.cuda())
    ref.backward()
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(world, port, kind, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        r = np.load(tmp_path / f"r{rank}.npz")
        assert float(r["loss"]) == pytest.approx(float(ref.detach()), rel=1e-5)
        lo, hi = pcl.shard_bounds(pred.shape[0], world, rank)
        np.testing.assert_allclose(r["grad"] / world, p.grad[lo:hi].cpu().numpy(), rtol=1e-5, atol=1e-10)
