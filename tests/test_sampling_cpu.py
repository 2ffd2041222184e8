"""CPU tests of the sampling oracles: the FPS oracle against the torch algorithm the
reference keeps in models/pointnet2_utils.py:64-86, the ball-query oracle against golden vectors produced by the
reference's own query_ball_point (tests/golden/make_golden.py)."""
import numpy as np
import torch

import oracle
from pointcloud_b200 import synth


def fps_torch(xyz, npoint):  # pointnet2_utils.py:64-86 with the deterministic start index 0
    B, N, _ = xyz.shape
    centroids = torch.zeros(B, npoint, dtype=torch.long)
    distance = torch.ones(B, N) * 1e10
    farthest = torch.zeros(B, dtype=torch.long)
    batch = torch.arange(B)
    for i in range(npoint):
        centroids[:, i] = farthest
        centroid = xyz[batch, farthest, :].view(B, 1, 3)
        dist = torch.sum((xyz - centroid) ** 2, -1)
        mask = dist < distance
        distance[mask] = dist[mask]
        farthest = torch.max(distance, -1)[1]
    return centroids


def test_fps_oracle_matches_reference_torch_algorithm():
    for b, n, m, seed in [(2, 500, 64, 0), (1, 2048, 512, 1), (3, 77, 77, 2)]:
        x, _ = synth.uniform_clouds(b, n, seed=seed)
        assert np.array_equal(oracle.fps(x, m), fps_torch(x, m).numpy())
    _, t = synth.table_clouds(2, 1024, seed=3)
    idx = oracle.fps(t[:, :, :3], 256)
    assert np.array_equal(idx, fps_torch(t[:, :, :3].contiguous(), 256).numpy())
    assert idx[:, 0].tolist() == [0, 0] and all(len(np.unique(r)) == 256 for r in idx)
    # duplicates: once every distinct location is taken the running minima are all 0 and index 0 repeats
    d = torch.tensor([[[0.1, 0.2, 0.3]] * 4 + [[0.9, 0.9, 0.9]]])
    assert oracle.fps(d, 4)[0].tolist() == [0, 4, 0, 0]
    # pointnet2_ops' padding convention: points at the origin are never selected
    z = torch.cat([torch.zeros(1, 3, 3), torch.rand(1, 20, 3, generator=torch.Generator().manual_seed(1)) + 0.2], dim=1)
    assert (oracle.fps(z, 8, skip_origin=True)[0, 1:] >= 3).all()


def test_ball_query_oracle_matches_reference_golden(golden):
    xyz, new_xyz = golden["bq_xyz"], golden["bq_new_xyz"]
    assert np.array_equal(oracle.fps(xyz, 128), golden["bq_fps_idx"])
    for name in "abc":
        radius, nsample = golden[f"bq_{name}_params"]
        got = oracle.ball_query(float(radius), int(nsample), xyz, new_xyz)
        ref, amb = golden[f"bq_{name}_idx"], golden[f"bq_{name}_ambiguous"]
        assert np.array_equal(got[~amb], ref[~amb])                      # exact away from the sphere's surface
        assert (got[amb] != ref[amb]).mean() < 0.2                        # and nearly everywhere on it
        own = np.take_along_axis(got, np.zeros_like(got[:, :, :1]), 2)[..., 0]
        assert (own <= golden["bq_fps_idx"]).all()                       # a centroid lies inside its own ball: the first hit is it or an earlier index
