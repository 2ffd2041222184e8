#!/usr/bin/env python
"""BASELINE config 4: one PointNet2-Autoencoder training step with the batch-sharded B200 loss.

This is a MEASUREMENT HARNESS, not part of the product: the reference's model zoo is out of scope, but
config 4 needs a producer of `pred`, so the architecture of `train.py:80-82` is re-stated compactly here
  encoder  = PointNet2Encoder (models/pointnet2.py:20-22): SA(512, r=0.2, 32, [64,64,128]) -> SA(128, r=0.4, 64,
             [128,128,256]) -> group-all [256,512,1024], then Linear(1024 -> bottleneck 13) (architectures.py:112-124)
  decoder  = MLP 13 -> [512,1024,2048] -> 2048*6, Sigmoid (architectures.py:141-155)
  step     = Lit.training_step + Adam(lr 1e-3) (train.py:30-35,67-68), 16-bit autocast (cfg.py:13)
with this repo's kernels in the three places the reference calls native code: farthest point sampling and ball query
inside sample_and_group (models/pointnet2_utils.py:116-144) and the EarthMoverDistance loss (utils.py:245-309), wrapped
in ShardedLoss when launched with torchrun.

    python examples/pointnet2_ae_step.py --steps 20
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/pointnet2_ae_step.py --steps 20
Prints one JSON line: step time, share of the loss in the step, clouds/s (whole job).
"""
import argparse
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pointcloud_b200 as pcl  # noqa: E402
from pointcloud_b200 import synth  # noqa: E402


def gather(points, idx):
    """points (B, N, C), idx (B, ...) -> (B, ..., C)"""
    b = points.shape[0]
    flat = idx.reshape(b, -1)
    out = torch.gather(points, 1, flat.unsqueeze(-1).expand(-1, -1, points.shape[2]))
    return out.reshape(*idx.shape, points.shape[2])


class SetAbstraction(nn.Module):
    def __init__(self, npoint, radius, nsample, in_channel, mlp):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        chans = [in_channel] + list(mlp)
        self.convs = nn.ModuleList(nn.Conv2d(a, b, 1) for a, b in zip(chans[:-1], chans[1:]))
        self.bns = nn.ModuleList(nn.BatchNorm2d(b) for b in chans[1:])

    def forward(self, xyz, feats):  # xyz (B, N, 3) fp32, feats (B, N, D) or None
        if self.npoint is None:  # group all
            new_xyz = xyz.new_zeros(xyz.shape[0], 1, 3)
            grouped = xyz.unsqueeze(1) if feats is None else torch.cat([xyz, feats], dim=-1).unsqueeze(1)
        else:
            with torch.no_grad():  # indices only: the two sm_100a sampling kernels
                fps_idx = pcl.farthest_point_sample(xyz, self.npoint)
                new_xyz = gather(xyz, fps_idx)
                idx = pcl.query_ball_point(self.radius, self.nsample, xyz, new_xyz)
            grouped = gather(xyz, idx) - new_xyz.unsqueeze(2)
            if feats is not None:
                grouped = torch.cat([grouped, gather(feats, idx)], dim=-1)
        x = grouped.permute(0, 3, 2, 1)  # (B, C, nsample, npoint)
        for conv, bn in zip(self.convs, self.bns):
            x = F.relu(bn(conv(x)))
        return new_xyz, x.max(dim=2)[0].permute(0, 2, 1)  # (B, npoint, C')


class PointNet2AE(nn.Module):
    def __init__(self, out_points=2048, out_dim=6, bottleneck=13):
        super().__init__()
        self.sa1 = SetAbstraction(512, 0.2, 32, 3 + 3, [64, 64, 128])
        self.sa2 = SetAbstraction(128, 0.4, 64, 128 + 3, [128, 128, 256])
        self.sa3 = SetAbstraction(None, None, None, 256 + 3, [256, 512, 1024])
        self.bottleneck = nn.Linear(1024, bottleneck)
        self.decoder = nn.Sequential(nn.Linear(bottleneck, 512), nn.ReLU(), nn.Linear(512, 1024), nn.ReLU(),
                                     nn.Linear(1024, 2048), nn.ReLU(), nn.Linear(2048, out_points * out_dim), nn.Sigmoid(),
                                     nn.Unflatten(1, (out_points, out_dim)))

    def forward(self, cloud):  # (B, N, 6) xyz + rgb
        xyz, rgb = cloud[:, :, :3].float().contiguous(), cloud[:, :, 3:]
        x1, f1 = self.sa1(xyz, rgb)
        x2, f2 = self.sa2(x1, f1)
        _, f3 = self.sa3(x2, f2)
        return self.decoder(self.bottleneck(f3.reshape(f3.shape[0], 1024)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch-per-gpu", type=int, default=32)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = PointNet2AE().to(dev)
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    loss_fn = pcl.EarthMoverDistance(eps=pcl.cfg.emd_eps, its=pcl.cfg.emd_iterations)
    if world > 1:
        loss_fn = pcl.ShardedLoss(loss_fn)
    _, target = synth.autoencoder_batch(args.batch_per_gpu, 2048, seed=rank, regime="independent")  # the AE reconstructs its input
    target = target.to(dev)
    amp = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": None}[args.dtype]
    scaler = torch.amp.GradScaler("cuda", enabled=(args.dtype == "fp16"))
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t_step, t_loss, losses = [], [], []
    for it in range(args.warmup + args.steps):
        a, b, c, d = ev(), ev(), ev(), ev()
        a.record()
        with torch.autocast("cuda", dtype=amp, enabled=amp is not None):
            pred = model(target)
        b.record()
        loss = loss_fn(pred, target)  # train.py:33 -- pred arrives in the autocast dtype, the kernels up-cast
        c.record()
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        d.record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            t_step.append(a.elapsed_time(d)); t_loss.append(b.elapsed_time(c)); losses.append(float(loss.detach()))
    ms = sum(t_step) / len(t_step)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t)
    if rank == 0:
        print(json.dumps({"config": "4: PointNet2 AE training step, EMD loss (eps 0.005, 50 iters), Adam, autocast " + args.dtype,
                          "n_gpus": world, "batch_per_gpu": args.batch_per_gpu, "ms_per_step": ms, "clouds_per_s": world * args.batch_per_gpu / (ms * 1e-3),
                          "loss_forward_ms": sum(t_loss) / len(t_loss), "loss_forward_share": sum(t_loss) / sum(t_step),
                          "first_loss": losses[0], "last_loss": losses[-1], "sharded": world > 1,
                          "params_M": sum(p.numel() for p in model.parameters()) / 1e6}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
