"""torch-CPU restatement of the reference's loss callables.  TEST INFRASTRUCTURE ONLY.

Follows pointcloud_vision/utils.py:207-309 line by line, with the native pieces replaced by the C
oracles of this package (emd_oracle.c for `emdModule`, chamfer_oracle.c for pytorch3d's knn).  It is
pinned against outputs of the REAL reference Python (imported with stub native modules) by
tests/golden/make_golden.py -> tests/golden/loss_golden.npz -> tests/test_oracle_cpu.py.
"""
from functools import reduce

import numpy as np
import torch
import torch.nn.functional as F

import oracle


class _EmdFn(torch.autograd.Function):
    """emdFunction (loss/emd/emd_module.py:31-72) on the CPU oracle."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, eps, iters):
        assert xyz1.shape[1] == xyz2.shape[1] and xyz1.shape[0] == xyz2.shape[0]
        x1, x2 = xyz1.detach().contiguous().float(), xyz2.detach().contiguous().float()
        r = oracle.emd_forward(x1, x2, eps, iters, nthreads=8)
        dist, asg = torch.from_numpy(r["dist"]), torch.from_numpy(r["assignment"])
        ctx.save_for_backward(x1, x2, asg)
        ctx.mark_non_differentiable(asg)
        ctx.stats = r
        return dist, asg

    @staticmethod
    def backward(ctx, graddist, gradidx):
        x1, x2, asg = ctx.saved_tensors
        g1, g2 = oracle.emd_backward(x1, x2, asg.numpy(), graddist.contiguous().float())
        return torch.from_numpy(g1), torch.from_numpy(g2), None, None


def emd(xyz1, xyz2, eps, iters):
    return _EmdFn.apply(xyz1, xyz2, eps, iters)


class _ChamferFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, x_lengths, y_lengths, mode):
        xc, yc = x.detach().contiguous().float(), y.detach().contiguous().float()
        r = oracle.chamfer_forward(xc, yc, x_lengths, y_lengths, mode=mode, nthreads=8)
        ctx.save_for_backward(xc, yc)
        ctx.aux = (r["idx_x"], r["idx_y"], x_lengths, y_lengths)
        ctx.raw = r
        return torch.tensor(float(r["loss"]), dtype=torch.float32)

    @staticmethod
    def backward(ctx, g):
        xc, yc = ctx.saved_tensors
        ix, iy, xl, yl = ctx.aux
        gx, gy = oracle.chamfer_backward(xc, yc, ix, iy, float(g), xl, yl)
        return torch.from_numpy(gx), torch.from_numpy(gy), None, None, None


def chamfer_distance(x, y, x_lengths=None, y_lengths=None, mode=0):
    """pytorch3d.loss.chamfer_distance(x, y, x_lengths, y_lengths) -> (loss, None) with the defaults (App. B)."""
    return _ChamferFn.apply(x, y, x_lengths, y_lengths, mode), None


class FilterClasses:  # utils.py:110-124
    def __init__(self, whitelist, label_dim):
        self.whitelist, self.label_dim = whitelist, label_dim

    def __call__(self, points):
        label = points[:, self.label_dim].long()
        mask = reduce(torch.logical_or, [label == v for v in self.whitelist])
        return points[mask, :]


class ChamferDistance:  # utils.py:209-211
    def __init__(self, mode=0):
        self.mode = mode

    def __call__(self, pred, target):
        return chamfer_distance(pred, target, mode=self.mode)[0]


class FilteringChamferDistance:  # utils.py:213-228
    def __init__(self, filter, mode=0):
        self.filter, self.mode = filter, mode

    def __call__(self, pred, target):
        pred = pred.to(dtype=torch.float32)
        filtered = [self.filter(p)[:, :3] for p in target]
        num_points = [p.shape[0] for p in filtered]
        max_points = max(num_points)
        target = torch.stack([F.pad(p, (0, 0, 0, max_points - p.shape[0])) for p in filtered]).to(dtype=torch.float32)
        return chamfer_distance(pred, target, y_lengths=torch.tensor(num_points), mode=self.mode)[0]


class SegmentingChamferDistance:  # utils.py:230-243
    def __init__(self, class_labels, mode=0):
        self.classs_losses = {c: FilteringChamferDistance(FilterClasses([l], label_dim=3), mode) for c, l in class_labels.items()}

    def __call__(self, pred, target):
        return torch.stack([loss(pred[c], target) for c, loss in self.classs_losses.items()]).sum()


class EarthMoverDistance:  # utils.py:245-309
    def __init__(self, eps=0.002, its=10000, num_classes=None, feature_weight=0.1, emd_fn=None):
        """`emd_fn(xyz1, xyz2, eps, iters) -> (dist, assignment)`: default = the CPU oracle; the GPU tests pass the CUDA
        emdModule to check it under the reference's own op-by-op loss structure."""
        self.eps, self.iterations, self.C, self.feature_weight = eps, its, num_classes, feature_weight
        self.emd_fn = emd_fn or emd
        self.logged = {}

    def log(self, name, value):
        self.logged[name] = float(value)

    def __call__(self, pred, target):
        dists, assignment = self.emd_fn(pred[:, :, :3], target[:, :, :3], self.eps, self.iterations)
        assignment = assignment.long().unsqueeze(-1)
        target = target.to(dists.device).take_along_dim(assignment, 1)
        pred = pred.to(dists.device)
        weights = torch.ones_like(dists)
        if self.C is not None:
            target_classes = target[:, :, 3].long()
            distribution = torch.bincount(target_classes.view(-1), minlength=self.C)
            distribution = distribution / distribution.sum()
            pred_classes = pred[:, :, 3:].argmax(dim=2)
            pred_distribution = torch.bincount(pred_classes.view(-1), minlength=self.C)
            pred_distribution = pred_distribution / pred_distribution.sum()
            kl_div = F.kl_div(F.log_softmax(pred_distribution, dim=0), F.softmax(distribution, dim=0), reduction='batchmean')
            class_weights = (1 / (distribution + 1e-4)) ** (1 - 0)
            class_weights = class_weights / class_weights.sum()
            weights = class_weights[target_classes]
            ce_l = F.cross_entropy(pred.permute(0, 2, 1)[:, 3:, :].float(), target_classes, weight=class_weights)
            feature_l = 0.1 * ce_l
            self.log('train_loss/cross_entropy', ce_l)
            self.log('train_loss/kl_divergence', kl_div)
        else:
            feature_l = F.mse_loss(pred[:, :, 3:], target[:, :, 3:])
        point_l = (dists.sqrt() * weights).sum() / weights.sum()
        self.log('train_loss/EMD', point_l)
        self.log('train_loss/feature', feature_l)
        return point_l + feature_l


def to_np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
