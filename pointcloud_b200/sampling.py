"""Farthest point sampling and ball query -- B200 kernels behind the names the reference uses on the producer side
of the loss path: `farthest_point_sample` / `query_ball_point` (models/pointnet2_utils.py:89-113) and
`sample_farthest_points` (pytorch3d call surface used at utils.py:90 and models/pointmlp.py:158)."""
import torch

from . import _lib


def farthest_point_sample(xyz, npoint, start_idx=None, skip_origin=True):
    """xyz (B, N, 3+) -> sampled point indices (B, npoint), int64 like the reference wrapper (pointnet2_utils.py:89-90).

    `skip_origin=True` (default) is the behaviour of the kernel that wrapper calls (pointnet2_ops `furthest_point_sample`
    never selects a point with x^2+y^2+z^2 <= 1e-3 after the first [3P-memory]); `skip_origin=False` is the plain torch
    algorithm kept as a comment in pointnet2_utils.py:64-86 and pytorch3d's `sample_farthest_points`."""
    _lib.require_cuda()
    L = _lib.lib()
    xyz = _lib.as_points(xyz)
    b, n, c = xyz.shape
    assert c >= 3
    dev = xyz.device
    st = None if start_idx is None else start_idx.to(device=dev, dtype=torch.int32).contiguous()
    with _lib.on_device(dev):
        idx = torch.empty(b, int(npoint), device=dev, dtype=torch.int32)
        rc = L.pcl_fps(*_lib.pts_args(xyz), b, n, int(npoint), _lib.ptr(st), 1 if skip_origin else 0, idx.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(rc, "pcl_fps")
    return idx.long()


def sample_farthest_points(points, lengths=None, K=50, random_start_point=False):
    """pytorch3d.ops.sample_farthest_points(points, K=K) -> (sampled points (B, K, D), indices (B, K)) for the call the
    reference makes (utils.py:90: one cloud, K points, all D channels returned, distances on xyz... pytorch3d uses ALL
    D channels for the distance; the reference only passes xyz+rgb rows through it in the dataset transform)."""
    if lengths is not None:
        raise NotImplementedError("lengths is not used by the reference (utils.py:90, pointmlp.py:158)")
    start = None
    if random_start_point:
        start = torch.randint(0, points.shape[1], (points.shape[0],))
    idx = farthest_point_sample(points[:, :, :3], K, start_idx=start, skip_origin=False)
    gathered = torch.gather(points.to(idx.device), 1, idx.unsqueeze(-1).expand(-1, -1, points.shape[2]))
    return gathered, idx


def query_ball_point(radius, nsample, xyz, new_xyz):
    """pointnet2_utils.py:93-113: (B, S, nsample) int64 indices of the first nsample points within `radius` of every
    centroid, padded with the first hit."""
    _lib.require_cuda()
    L = _lib.lib()
    xyz, new_xyz = _lib.as_points(xyz), _lib.as_points(new_xyz)
    b, n, _ = xyz.shape
    s = new_xyz.shape[1]
    dev = xyz.device
    r2 = float(torch.tensor(radius ** 2, dtype=torch.float32))  # the comparison happens in fp32 (sqrdists > radius ** 2)
    with _lib.on_device(dev):
        out = torch.empty(b, s, int(nsample), device=dev, dtype=torch.int32)
        rc = L.pcl_ball_query(*_lib.pts_args(xyz), *_lib.pts_args(new_xyz), b, n, s, r2, int(nsample), out.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(rc, "pcl_ball_query")
    return out.long()
