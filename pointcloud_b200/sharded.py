"""Batch-sharded loss wrapper: one shard of clouds per rank, no exchange of points.

Every cloud pair is independent in Chamfer, in the auction and in both backward passes, so rank r simply
evaluates the loss kernels on its own clouds.  The only cross-rank quantities are the whole-batch
statistics the reference's loss uses (utils.py:274-275 class histogram, :304 weights.sum(), the batch
mean of Chamfer): they are summed with `torch.distributed.all_reduce` (NCCL over NVLink on the GPU
box, gloo in the CPU tests) -- a handful of scalars per loss call.

Value and gradient contract
  * the returned tensor holds the GLOBAL loss (identical on every rank, equal to what one GPU would
    compute on the concatenated batch);
  * its backward yields `scale * d(global loss)/d(local pred)`; with `ddp_average=True` (default)
    scale = world_size, which cancels DistributedDataParallel's gradient averaging, so a DDP-wrapped
    model trained through this wrapper takes exactly the single-GPU step.
"""
import torch
import torch.distributed as dist

from .losses import ChamferDistance, EarthMoverDistance, FilteringChamferDistance, SegmentingChamferDistance


def _all_reduce_sum(t, group):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class _GlobalRatio:
    """(sum_r num_r) / (sum_r den_r): value is global, gradient flows through the local numerator only
    (denominators on this path never depend on the prediction)."""

    def __init__(self, group, scale):
        self.group, self.scale = group, scale

    def __call__(self, num, den):
        packed = torch.stack([num.detach().float(), den.detach().float()])
        g = _all_reduce_sum(packed, self.group)
        gnum, gden = g[0], g[1]
        local = num / gden                          # d/d num_local of the global ratio
        value = (gnum / gden).detach()
        # value of the global ratio, gradient of scale * local
        return value + self.scale * (local - local.detach())


def shard_bounds(batch: int, world: int, rank: int):
    """Contiguous batch shards: rank r owns clouds [r*B/G, (r+1)*B/G)."""
    lo = (batch * rank) // world
    hi = (batch * (rank + 1)) // world
    return lo, hi


class ShardedLoss:
    """Wrap one of the loss callables so that `loss(pred_local, target_local)` evaluates the loss of the
    whole (sharded) batch.  Usable as `Lit.loss_fn` (train.py:33) under DDP without other changes."""

    def __init__(self, loss_fn, process_group=None, ddp_average=True):
        self.loss_fn = loss_fn
        self.group = process_group
        self.ddp_average = ddp_average
        if not isinstance(loss_fn, (EarthMoverDistance, ChamferDistance, FilteringChamferDistance, SegmentingChamferDistance)):
            raise TypeError(f"ShardedLoss does not know how to shard {type(loss_fn).__name__}")

    # the `.log` protocol of train.py:161 is forwarded to the wrapped loss
    @property
    def log(self):
        return self.loss_fn.log

    @log.setter
    def log(self, fn):
        self.loss_fn.log = fn

    def _world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group)
        return 1

    def __call__(self, pred, target):
        world = self._world()
        scale = float(world) if self.ddp_average else 1.0
        fn = self.loss_fn
        if isinstance(fn, EarthMoverDistance):
            if not fn.fused:
                raise ValueError("ShardedLoss needs EarthMoverDistance(fused=True)")
            fn.reduce_hist = lambda h: _all_reduce_sum(h, self.group)
            fn.reduce_ratio = _GlobalRatio(self.group, scale)
            try:
                return fn(pred, target)
            finally:
                fn.reduce_hist = None
                fn.reduce_ratio = None
        # Chamfer family: loss = sum over local clouds / B_global  (batch_reduction="mean")
        some = next(iter(pred.values())) if isinstance(pred, dict) else pred
        b_local = torch.tensor([float(some.shape[0])], device=some.device)
        b_global = _all_reduce_sum(b_local, self.group)[0]
        local = fn(pred, target) * (b_local[0] / b_global)     # local mean -> share of the global mean
        value = _all_reduce_sum(local.detach().reshape(1), self.group)[0]
        return value + scale * (local - local.detach())
