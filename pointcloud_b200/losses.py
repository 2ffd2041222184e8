"""Loss callables -- B200 drop-in for the loss section of pointcloud_vision/utils.py:207-309.

Same class names, constructor arguments, call signature `loss(pred, target) -> 0-d tensor` and the
optional `.log` attribute protocol (train.py:161 assigns `model.loss_fn.log = model.log`; the loss calls
`self.log(name, tensor)` with the keys of utils.py:297-298,306-307).
"""
from functools import reduce

import torch
import torch.nn.functional as F
from torch.autograd import Function

from . import _lib, cfg
from .chamfer import chamfer_distance
from .emd_module import emdModule, emd_forward_raw, emd_workspace


class FilterClasses:
    """utils.py:110-124: keep the points whose label (column `label_dim`) is in the whitelist."""

    def __init__(self, whitelist, label_dim):
        self.whitelist = whitelist
        self.label_dim = label_dim

    def __call__(self, points):
        label = points[:, self.label_dim].long()  # (N,)
        mask = reduce(torch.logical_or, [label == v for v in self.whitelist])
        return points[mask, :]


########## Loss Functions ##########

class ChamferDistance:
    """utils.py:209-211 (all feature channels take part in the distance)."""

    def __call__(self, pred, target):
        return chamfer_distance(pred, target)[0]


class FilteringChamferDistance:
    """utils.py:213-228: per-cloud class filter of the target, pad + stack, Chamfer with y_lengths.

    With a `FilterClasses` filter the per-cloud Python loop of the reference (utils.py:222-226, one
    boolean-index + host sync per cloud) is replaced by one batched stable compaction (argsort of the
    mask) and a single host read of max(num_points); any other callable falls back to the reference's
    loop."""

    def __init__(self, filter):
        self.filter = filter

    def _filter_pad(self, target, dtype):
        f = self.filter
        if (isinstance(f, FilterClasses) and target.dim() == 3 and target.is_cuda and dtype == torch.float32
                and target.dtype in (torch.float32, torch.float16, torch.bfloat16) and target.stride(2) == 1
                and 0 < len(f.whitelist) <= 16 and target.shape[1] > 0 and target.shape[2] >= 3
                and 0 <= int(f.label_dim) < target.shape[2] and not target.requires_grad):
            # one launch, no host synchronisation: kept points packed to the front of a (B, N, 3) buffer + their counts
            import ctypes
            L = _lib.lib()
            b, n, _ = target.shape
            with _lib.on_device(target.device):
                xyz = torch.empty(b, n, 3, device=target.device, dtype=torch.float32)
                num_points = torch.empty(b, device=target.device, dtype=torch.int64)
                labels = (ctypes.c_int64 * len(f.whitelist))(*[int(v) for v in f.whitelist])
                rc = L.pcl_class_filter(*_lib.pts_args(target), b, n, int(f.label_dim), labels, len(f.whitelist),
                                        xyz.data_ptr(), num_points.data_ptr(), _lib.stream_ptr(target.device))
                _lib.check(rc, "pcl_class_filter")
            return xyz, num_points
        if isinstance(f, FilterClasses) and target.dim() == 3:
            label = target[:, :, f.label_dim].long()                                   # (B, N)
            mask = reduce(torch.logical_or, [label == v for v in f.whitelist])         # (B, N)
            num_points = mask.sum(dim=1)                                               # (B,)
            max_points = int(num_points.max().item()) if mask.shape[0] > 0 else 0
            order = torch.argsort((~mask).to(torch.uint8), dim=1, stable=True)[:, :max_points]  # kept points first, in order
            xyz = target[:, :, :3].gather(1, order.unsqueeze(-1).expand(-1, -1, 3)).to(dtype=dtype)
            keep = torch.arange(max_points, device=target.device)[None, :] < num_points[:, None]
            xyz = xyz * keep.unsqueeze(-1).to(dtype)                                    # F.pad zeros (utils.py:226)
            return xyz, num_points
        filtered = [f(p)[:, :3] for p in target]
        num_points = [p.shape[0] for p in filtered]
        max_points = max(num_points)
        padded = torch.stack([F.pad(p, (0, 0, 0, max_points - p.shape[0])) for p in filtered]).to(dtype=dtype)
        return padded, torch.tensor(num_points, device=target.device)

    def __call__(self, pred, target):
        device, dtype = pred.device, torch.float32  # utils.py:218
        pred = pred.to(dtype=dtype)
        target, num_points = self._filter_pad(target, dtype)
        return chamfer_distance(pred, target, y_lengths=num_points.to(device))[0]


class SegmentingChamferDistance:
    """utils.py:230-243: sum over classes of FilteringChamferDistance(pred[class], target)."""

    def __init__(self, class_labels):
        self.classs_losses = {c: FilteringChamferDistance(FilterClasses([l], label_dim=3)) for c, l in class_labels.items()}

    def __call__(self, pred, target):
        loss_per_class = torch.stack([loss(pred[c], target) for c, loss in self.classs_losses.items()])
        return loss_per_class.sum()


class _MatchedPointLoss(Function):
    """point_l = sum(w * sqrt(dist)) / sum(w) with (dist, assignment) from the auction, fused
    (utils.py:254,292,304 + emd_module.py:63-72).  Returns the two sums so that a batch-sharded caller
    can all-reduce them before dividing."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, dist, assignment, matched, class_weights):
        L = _lib.lib()
        b, n = dist.shape
        dev = dist.device
        c = 0 if class_weights is None else class_weights.numel()
        with _lib.on_device(dev):
            sums = torch.empty(2, device=dev, dtype=torch.float32)
            ws, wsb = emd_workspace(b, n, dev)
            rc = L.pcl_emd_weighted_reduce(dist.data_ptr(), _lib.ptr(matched), _lib.ptr(class_weights), b, n, c,
                                           sums.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr(dev))
            _lib.check(rc, "pcl_emd_weighted_reduce")
        ctx.save_for_backward(xyz1, xyz2, dist, assignment, matched, class_weights)
        ctx.in_meta = (xyz1.dtype,)
        return sums

    @staticmethod
    def backward(ctx, grad_sums):
        # Only d/d sums[0] reaches the points (sums[1] depends on labels only).  The kernel computes
        # grad_xyz1 = g0 * w / (2 sqrt(dist)) * 2 (xyz1 - xyz2[assignment]) with g0 read on the device;
        # `unit` makes its 1/sums[1] factor a no-op so that the division stays in the autograd graph.
        xyz1, xyz2, dist, assignment, matched, class_weights = ctx.saved_tensors
        L = _lib.lib()
        b, n = dist.shape
        dev = dist.device
        c = 0 if class_weights is None else class_weights.numel()
        g = grad_sums.contiguous().float()
        with _lib.on_device(dev):
            unit = torch.ones(2, device=dev, dtype=torch.float32)
            grad = torch.empty(b, n, 3, device=dev, dtype=torch.float32)
            rc = L.pcl_emd_weighted_bwd(*_lib.pts_args(xyz1), *_lib.pts_args(xyz2), b, n, assignment.data_ptr(),
                                        dist.data_ptr(), _lib.ptr(matched), _lib.ptr(class_weights), c,
                                        unit.data_ptr(), g.data_ptr(), grad.data_ptr(), _lib.stream_ptr(dev))
            _lib.check(rc, "pcl_emd_weighted_bwd")
        if xyz1.shape[2] != 3:
            full = torch.zeros(xyz1.shape, device=dev, dtype=torch.float32)
            full[:, :, :3] = grad
            grad = full
        return grad.to(ctx.in_meta[0]), None, None, None, None, None


class _FusedPointLoss(Function):
    """Unweighted point term (utils.py:304 with weights == 1): the sums and d(sum sqrt(dist))/d xyz1 were written by the
    auction kernel's epilogue (pcl_emd_fwd_fused), so forward launches nothing and backward is one scale."""

    @staticmethod
    def forward(ctx, xyz1, unit_grad, sums):
        ctx.save_for_backward(unit_grad)
        ctx.in_meta = (xyz1.dtype, xyz1.shape)
        return sums[:2].clone()

    @staticmethod
    def backward(ctx, grad_sums):
        (unit_grad,) = ctx.saved_tensors
        dt, shape = ctx.in_meta
        grad = unit_grad * grad_sums[0]
        if shape[2] != 3:
            full = torch.zeros(shape, device=grad.device, dtype=torch.float32)
            full[:, :, :3] = grad
            grad = full
        return grad.to(dt), None, None


def matched_label_hist(target_label, assignment, num_classes):
    """(hist int64[C], matched int32 (B,N)): labels of the targets each prediction was matched to and
    their histogram (utils.py:257-258,271-275 -- the bincount of the PERMUTED target labels)."""
    L = _lib.lib()
    b, n = assignment.shape
    dev = assignment.device
    lab = _lib.as_points(target_label)
    with _lib.on_device(dev):
        hist = torch.empty(num_classes, device=dev, dtype=torch.int64)
        matched = torch.empty(b, n, device=dev, dtype=torch.int32)
        rc = L.pcl_emd_match_hist(*_lib.pts_args(lab), assignment.data_ptr(), b, n, num_classes, hist.data_ptr(),
                                  matched.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(rc, "pcl_emd_match_hist")
    return hist, matched


class _SegCrossEntropySums(Function):
    """(sum_i w_i nll_i, sum_i w_i) of the class-weighted cross entropy (utils.py:293-295) and the histogram of
    argmax(logits) (utils.py:278-279), one kernel; backward: w_i (softmax - onehot), one kernel."""

    @staticmethod
    def forward(ctx, logits, matched, class_weights):
        L = _lib.lib()
        b, n, c = logits.shape
        dev = logits.device
        with _lib.on_device(dev):
            sums = torch.empty(2, device=dev, dtype=torch.float32)
            pred_hist = torch.empty(c, device=dev, dtype=torch.int64)
            wsb = L.pcl_emd_feature_workspace_bytes()
            ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
            rc = L.pcl_emd_seg_ce_fwd(*_lib.pts_args(logits), matched.data_ptr(), class_weights.data_ptr(), b, n, c,
                                      sums.data_ptr(), pred_hist.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr(dev))
            _lib.check(rc, "pcl_emd_seg_ce_fwd")
        ctx.save_for_backward(logits, matched, class_weights)
        ctx.mark_non_differentiable(pred_hist)
        return sums, pred_hist

    @staticmethod
    def backward(ctx, grad_sums, _grad_hist):
        logits, matched, class_weights = ctx.saved_tensors
        L = _lib.lib()
        b, n, c = logits.shape
        dev = logits.device
        g = grad_sums.contiguous().float()
        with _lib.on_device(dev):
            grad = torch.empty(b, n, c, device=dev, dtype=torch.float32)
            rc = L.pcl_emd_seg_ce_bwd(*_lib.pts_args(logits), matched.data_ptr(), class_weights.data_ptr(), b, n, c,
                                      g.data_ptr(), grad.data_ptr(), _lib.stream_ptr(dev))
            _lib.check(rc, "pcl_emd_seg_ce_bwd")
        return grad.to(logits.dtype), None, None


class _MatchedFeatureMSESums(Function):
    """(sum (pred_feat - target_feat[assignment])^2, element count): F.mse_loss against the permuted target
    (utils.py:257-258,301) without materialising the permutation; backward: 2 (pred_feat - target_feat[assignment])."""

    @staticmethod
    def forward(ctx, feat, tfeat, assignment):
        L = _lib.lib()
        b, n, f = feat.shape
        dev = feat.device
        with _lib.on_device(dev):
            sums = torch.empty(2, device=dev, dtype=torch.float32)
            wsb = L.pcl_emd_feature_workspace_bytes()
            ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
            rc = L.pcl_emd_feat_mse_fwd(*_lib.pts_args(feat), *_lib.pts_args(tfeat), assignment.data_ptr(), b, n, f,
                                        sums.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr(dev))
            _lib.check(rc, "pcl_emd_feat_mse_fwd")
        ctx.save_for_backward(feat, tfeat, assignment)
        return sums

    @staticmethod
    def backward(ctx, grad_sums):
        feat, tfeat, assignment = ctx.saved_tensors
        L = _lib.lib()
        b, n, f = feat.shape
        dev = feat.device
        g = grad_sums.contiguous().float()
        with _lib.on_device(dev):
            grad = torch.empty(b, n, f, device=dev, dtype=torch.float32)
            rc = L.pcl_emd_feat_mse_bwd(*_lib.pts_args(feat), *_lib.pts_args(tfeat), assignment.data_ptr(), b, n, f,
                                        g.data_ptr(), grad.data_ptr(), _lib.stream_ptr(dev))
            _lib.check(rc, "pcl_emd_feat_mse_bwd")
        return grad.to(feat.dtype), None, None


class EarthMoverDistance:
    """utils.py:245-309.  The auction runs in one kernel; the loss around it (matched-label histogram, class
    weights, weighted sqrt-sum, weighted cross entropy / feature MSE and their backward passes) runs in the
    fused epilogue kernels instead of the reference's ~25 torch ops.

    `all_reduce` / `grad_scale` are the hook of the batch-sharded wrapper (sharded.py): `all_reduce` sums a
    small vector of whole-batch statistics over the ranks -- once for the class histogram (utils.py:274-275)
    and once for ALL numerators / denominators / the argmax histogram packed into one vector -- i.e. two
    dependent collectives for the Segmenter loss and one for the Autoencoder loss; `grad_scale` multiplies the
    gradient (DDP averaging).  The defaults keep single-GPU semantics without any extra op."""

    def __init__(self, eps=0.002, its=10000, num_classes=None, feature_weight=0.1):
        self.loss_fn = emdModule()
        self.eps = eps
        self.iterations = its
        self.C = num_classes
        self.feature_weight = feature_weight  # stored but unused, like the reference (utils.py:251,296)
        self.all_reduce = None   # set by ShardedLoss: tensor -> tensor summed over the ranks
        self.grad_scale = 1.0    # set by ShardedLoss

    def log(self, name, value):  # replaced by train.py:161 (`model.loss_fn.log = model.log`)
        pass

    # -- helpers shared by both paths ---------------------------------------------------------------
    def _class_weights(self, hist):
        if self.all_reduce is not None:
            hist = self.all_reduce(hist)                              # collective 1 of 2 (Segmenter only)
        distribution = hist / hist.sum()                              # utils.py:274-275
        class_weights = (1 / (distribution + 1e-4)) ** (1 - 0)        # :285
        class_weights = class_weights / class_weights.sum()           # :286
        return distribution, class_weights

    def _ratios(self, pairs, counts=None):
        """[num / den for (num, den) in pairs] (+ `counts`) over the WHOLE batch.  Single GPU: plain divisions.
        Sharded: every numerator, denominator and count travels in ONE all-reduced fp64 vector; the value is the
        global ratio, the gradient flows through the local numerator only (no denominator on this path depends
        on the prediction) and is multiplied by `grad_scale`."""
        if self.all_reduce is None:
            return [n / d for n, d in pairs], counts
        flat = [x.detach().double() for nd in pairs for x in nd]
        packed = torch.stack(flat) if counts is None else torch.cat([torch.stack(flat), counts.double()])
        g = self.all_reduce(packed)
        out = []
        for i, (n, _) in enumerate(pairs):
            gnum, gden = g[2 * i].float(), g[2 * i + 1].float()
            local = n / gden
            out.append((gnum / gden).detach() + self.grad_scale * (local - local.detach()))
        return out, (None if counts is None else g[2 * len(pairs):])

    def _kl(self, pred_hist, distribution):
        pred_distribution = pred_hist / pred_hist.sum()                              # utils.py:280
        return F.kl_div(F.log_softmax(pred_distribution.float(), dim=0), F.softmax(distribution, dim=0), reduction='batchmean')  # :283

    # -- the places where the fused path enters the CUDA library (overridden only by CPU host-logic tests)
    def _auction(self, pred, target, want_epilogue=False):
        """(xyz1, xyz2, dists, assignment[, unit_grad, sums]): the last two come from the kernel's fused epilogue and are
        only asked for by the unweighted loss; an override may ignore the flag and return four values."""
        xyz1, xyz2 = _lib.as_points(pred[:, :, :3]), _lib.as_points(target[:, :, :3])
        r = emd_forward_raw(xyz1, xyz2, self.eps, self.iterations, want_epilogue=want_epilogue)
        return (xyz1, xyz2, r[0], r[1]) + tuple(r[3:])

    def _matched_hist(self, target, assignment):
        return matched_label_hist(target[:, :, 3:4], assignment, self.C)

    def _point_sums(self, xyz1, xyz2, dists, assignment, matched, class_weights):
        return _MatchedPointLoss.apply(xyz1, xyz2, dists, assignment, matched, class_weights)

    def _ce_sums(self, pred, matched, class_weights):
        """(sum w*nll, sum w) of the weighted cross entropy and the argmax histogram of the logits (utils.py:278-279,293-295)."""
        return _SegCrossEntropySums.apply(_lib.as_points(pred[:, :, 3:]), matched, class_weights)

    def _mse_sums(self, pred, target, assignment):
        """(sum of squared feature differences against the permuted target, element count) (utils.py:257-258,301)."""
        return _MatchedFeatureMSESums.apply(_lib.as_points(pred[:, :, 3:]), _lib.as_points(target[:, :, 3:]), assignment)

    def __call__(self, pred, target):
        if not pred.is_cuda and type(self)._auction is EarthMoverDistance._auction:
            _lib.require_cuda()
            pred, target = pred.cuda(), target.cuda()
        res = self._auction(pred, target, want_epilogue=(self.C is None))
        xyz1, xyz2, dists, assignment = res[:4]
        epilogue = res[4:] if len(res) == 6 else None   # (unit_grad, sums) of the fused kernel epilogue

        if cfg.debug:  # utils.py:261-265
            num_points = pred.shape[1]
            num_missing = num_points - assignment.unique().numel()
            if num_missing / num_points > 0.005:
                print(f"DEBUG: EMD unassigned = {num_missing} / {num_points} = {num_missing / num_points}")

        if self.C is not None:  # segmentation (utils.py:269-298)
            hist, matched = self._matched_hist(target, assignment)
            distribution, class_weights = self._class_weights(hist)
            class_weights = class_weights.float().contiguous()
            # weighted cross entropy == sum_i w_i * nll_i / sum_i w_i  (F.cross_entropy with weight=, utils.py:295)
            ce_sums, pred_hist = self._ce_sums(pred, matched, class_weights)
            sums = self._point_sums(xyz1, xyz2, dists, assignment, matched, class_weights)
            (ce_l, point_l), pred_hist = self._ratios([(ce_sums[0], ce_sums[1]), (sums[0], sums[1])], pred_hist)  # collective 2 of 2
            kl_div = self._kl(pred_hist, distribution)
            feature_l = 0.1 * ce_l
            self.log('train_loss/cross_entropy', ce_l)
            self.log('train_loss/kl_divergence', kl_div)
        else:  # general feature loss (utils.py:300-301)
            if epilogue is not None:
                sums = _FusedPointLoss.apply(xyz1, epilogue[0], epilogue[1])
            else:
                sums = self._point_sums(xyz1, xyz2, dists, assignment, None, None)
            if pred.shape[2] == 3 or pred.numel() == 0:
                matched_feat = target[:, :, 3:].take_along_dim(assignment.long().unsqueeze(-1), 1)
                feature_l = F.mse_loss(pred[:, :, 3:], matched_feat)  # nan, exactly like the reference on empty features
                (point_l,), _ = self._ratios([(sums[0], sums[1])])
            else:
                mse_sums = self._mse_sums(pred, target, assignment)
                (feature_l, point_l), _ = self._ratios([(mse_sums[0], mse_sums[1]), (sums[0], sums[1])])  # the one collective

        # point_l: utils.py:304
        self.log('train_loss/EMD', point_l)
        self.log('train_loss/feature', feature_l)
        return point_l + feature_l


class _ChamferEmdStep(torch.autograd.Function):
    """Chamfer loss and unweighted EMD loss of one batch in ONE C-ABI call (`pcl_chamfer_emd_step`, include/pcl.h): the auction with
    its fused epilogue on the caller's stream, Chamfer forward + backward next to it on the library's side stream.  Both gradients are
    produced by the forward call (as the reference's training step needs them anyway); backward only scales and adds them."""

    @staticmethod
    def forward(ctx, pred, target, eps, iters, chamfer_mode):
        _lib.require_cuda()
        L = _lib.lib()
        p, t = _lib.as_points(pred), _lib.as_points(target)
        b, n, _ = p.shape
        assert t.shape[0] == b and t.shape[1] == n and p.shape[2] >= 3 and t.shape[2] >= 3
        dev = p.device
        with _lib.on_device(dev):
            out = torch.empty(8, device=dev, dtype=torch.float32)
            g_ch = torch.empty(b, n, 3, device=dev, dtype=torch.float32)
            g_emd = torch.empty(b, n, 3, device=dev, dtype=torch.float32)
            nbytes = _STEP_BYTES.get((b, n))
            if nbytes is None:
                nbytes = _STEP_BYTES[(b, n)] = L.pcl_chamfer_emd_step_scratch_bytes(b, n)
            scratch = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            rc = L.pcl_chamfer_emd_step(*_lib.pts_args(p), *_lib.pts_args(t), b, n, float(eps), int(iters), int(chamfer_mode), out.data_ptr(),
                                        g_ch.data_ptr(), g_emd.data_ptr(), scratch.data_ptr(), nbytes, _lib.stream_ptr(dev))
            _lib.check(rc, "pcl_chamfer_emd_step")
        ctx.save_for_backward(g_ch, g_emd)
        ctx.meta = (pred.shape, pred.dtype)
        return out[0] + out[1], out[6]

    @staticmethod
    def backward(ctx, g_chamfer, g_emd_loss):
        g_ch, g_emd = ctx.saved_tensors
        shape, dtype = ctx.meta
        g = torch.addcmul(g_ch * g_chamfer, g_emd, g_emd_loss)
        if shape[2] != 3:  # more channels than xyz: only xyz receives gradient
            full = torch.zeros(shape, device=g.device, dtype=torch.float32)
            full[:, :, :3] = g
            g = full
        return g.to(dtype), None, None, None, None


_STEP_BYTES = {}


def chamfer_emd_loss(pred, target, eps=0.005, iters=50, chamfer_mode=None):
    """(chamfer_loss, emd_loss) of a batch -- pytorch3d-style Chamfer (point mean, batch mean, both directions summed; utils.py:211) and
    mean sqrt of the auction distance (utils.py:304 with weights == 1) -- from ONE fused call.  Differentiable w.r.t. `pred` (B,N,>=3);
    `target` (B,N,>=3) gets no gradient.  Equivalent to chamfer_distance(pred, target)[0] and emdModule()(pred, target, eps, iters)[0]
    .sqrt().mean(), but the two losses run concurrently and their backward is a scale-and-add."""
    if chamfer_mode is None:
        chamfer_mode = 1 if cfg.chamfer_mode == "fma" else 0
    return _ChamferEmdStep.apply(pred, target, eps, iters, chamfer_mode)
