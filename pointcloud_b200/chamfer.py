"""Chamfer distance -- B200 drop-in for the subset of `pytorch3d.loss.chamfer_distance` the reference
uses (pointcloud_vision/utils.py:211 `chamfer_distance(pred, target)[0]` and :228
`chamfer_distance(pred, target, y_lengths=...)[0]`).

Semantics restated from pytorch3d 0.7.2 (parity unpinned -- pytorch3d is not available):
squared-L2 K=1 nearest neighbour in both directions (lowest index on exact ties), padded points ignored,
point_reduction="mean", batch_reduction="mean", loss = cham_x + cham_y; both clouds receive gradient.
"""
import torch
from torch.autograd import Function

from . import _lib, cfg

_MODES = {"unfused": _lib.PCL_CHAMFER_UNFUSED, "fma": _lib.PCL_CHAMFER_FMA, 0: 0, 1: 1}


def _lengths(lengths, b, p, dev, name):
    if lengths is None:
        return None
    if lengths.shape != (b,):
        raise ValueError(f"Expected {name} to be of shape (N,).")  # pytorch3d's message
    return lengths.to(device=dev, dtype=torch.int64).contiguous()


def chamfer_forward_raw(x, y, x_lengths=None, y_lengths=None, mode=None):
    """One pcl_chamfer_fwd call.  Returns dict(loss_xy (2,) batch means, loss_sums (2,) batch sums, dist_x, idx_x, dist_y, idx_y)."""
    _lib.require_cuda()
    L = _lib.lib()
    x, y = _lib.as_points(x), _lib.as_points(y)
    if y.device != x.device:
        y = y.to(x.device)
    b, p1, d = x.shape
    if y.shape[0] != b or y.shape[2] != d:
        raise ValueError("y does not have the correct shape.")  # pytorch3d's message
    p2 = y.shape[1]
    dev = x.device
    mode = _MODES[cfg.chamfer_mode if mode is None else mode]
    xl, yl = _lengths(x_lengths, b, p1, dev, "x_lengths"), _lengths(y_lengths, b, p2, dev, "y_lengths")
    with _lib.on_device(dev):
        dist_x = torch.empty(b, p1, device=dev, dtype=torch.float32); idx_x = torch.empty(b, p1, device=dev, dtype=torch.int32)
        dist_y = torch.empty(b, p2, device=dev, dtype=torch.float32); idx_y = torch.empty(b, p2, device=dev, dtype=torch.int32)
        loss4 = torch.empty(4, device=dev, dtype=torch.float32)
        wsb = L.pcl_chamfer_workspace_bytes(b, p1, p2)
        ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
        rc = L.pcl_chamfer_fwd(*_lib.pts_args(x), _lib.ptr(xl), *_lib.pts_args(y), _lib.ptr(yl), b, p1, p2, d, mode,
                               dist_x.data_ptr(), idx_x.data_ptr(), dist_y.data_ptr(), idx_y.data_ptr(),
                               loss4.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr(dev))
        _lib.check(rc, "pcl_chamfer_fwd")
    return dict(loss_xy=loss4[:2], loss_sums=loss4[2:], dist_x=dist_x, idx_x=idx_x, dist_y=dist_y, idx_y=idx_y, x=x, y=y, x_len=xl, y_len=yl)


class _ChamferFunction(Function):
    @staticmethod
    def forward(ctx, x, y, x_lengths, y_lengths, mode):
        ctx.in_meta = (x.dtype, x.device, y.dtype, y.device)
        r = chamfer_forward_raw(x, y, x_lengths, y_lengths, mode)
        ctx.save_for_backward(r["x"], r["y"], r["idx_x"], r["idx_y"])
        ctx.lens = (r["x_len"], r["y_len"])
        return r["loss_xy"]

    @staticmethod
    def backward(ctx, grad_xy):
        x, y, idx_x, idx_y = ctx.saved_tensors
        xl, yl = ctx.lens
        L = _lib.lib()
        b, p1, d = x.shape
        p2 = y.shape[1]
        dev = x.device
        grad_xy = grad_xy.contiguous().float()
        with _lib.on_device(dev):
            gx = torch.empty(b, p1, d, device=dev, dtype=torch.float32)
            gy = torch.empty(b, p2, d, device=dev, dtype=torch.float32)
            g = grad_xy  # (2,): upstream gradients of loss_x and loss_y, read on the device
            rc = L.pcl_chamfer_bwd(*_lib.pts_args(x), _lib.ptr(xl), *_lib.pts_args(y), _lib.ptr(yl), b, p1, p2, d,
                                   idx_x.data_ptr(), idx_y.data_ptr(), g.data_ptr(), gx.data_ptr(), gy.data_ptr(),
                                   _lib.stream_ptr(dev))
            _lib.check(rc, "pcl_chamfer_bwd")
        dtx, devx, dty, devy = ctx.in_meta
        return gx.to(device=devx, dtype=dtx), gy.to(device=devy, dtype=dty), None, None, None


def chamfer_distance(x, y, x_lengths=None, y_lengths=None, x_normals=None, y_normals=None, weights=None,
                     batch_reduction="mean", point_reduction="mean", norm=2, *, mode=None):
    """Same call surface as pytorch3d.loss.chamfer_distance for the arguments the reference uses
    (utils.py:211,228).  Returns (loss, None)."""
    if x_normals is not None or y_normals is not None or weights is not None:
        raise NotImplementedError("normals / weights are not on the reference's path (utils.py:211,228)")
    if batch_reduction != "mean" or point_reduction != "mean" or norm != 2:
        raise NotImplementedError("only batch_reduction='mean', point_reduction='mean', norm=2 (the defaults the reference uses)")
    if x.dim() != 3 or y.dim() != 3:
        raise ValueError("Expected points to be of shape (N, P, D)")
    loss_xy = _ChamferFunction.apply(x, y, x_lengths, y_lengths, mode)
    return loss_xy[0] + loss_xy[1], None
