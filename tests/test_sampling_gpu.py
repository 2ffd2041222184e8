"""GPU parity tests of farthest point sampling and ball query against the oracles, the reference golden
vectors, and size-independent properties."""
import numpy as np
import pytest
import torch

import oracle
import pointcloud_b200 as pcl
from pointcloud_b200 import synth
from helpers import npy

pytestmark = pytest.mark.gpu


def fps_plain(*a, **k):
    """The torch algorithm of pointnet2_utils.py:64-86 (what the oracle and the golden vectors follow): no point is skipped."""
    return pcl.farthest_point_sample(*a, skip_origin=False, **k)


@pytest.mark.parametrize("b,n,m", [(2, 500, 64), (4, 2048, 512), (1, 1, 1), (3, 77, 77), (2, 4096, 128), (1, 16384, 64), (2, 1000, 1)])
def test_fps_bit_exact_vs_oracle(b, n, m):
    x, _ = synth.uniform_clouds(b, n, seed=n + m)
    got = fps_plain(x.cuda(), m)
    assert got.dtype == torch.int64 and got.shape == (b, m)
    assert np.array_equal(npy(got), oracle.fps(x, m))


def test_fps_table_clouds_strides_dtypes_and_options():
    pred, target = synth.autoencoder_batch(4, 2048, seed=1)
    want = oracle.fps(target[:, :, :3], 512)
    assert np.array_equal(npy(fps_plain(target.cuda()[:, :, :3], 512)), want)          # row stride 6, no copy
    assert np.array_equal(npy(fps_plain(target.cuda(), 512)), want)                    # extra channels ignored
    h = target.cuda().half()
    assert np.array_equal(npy(fps_plain(h[:, :, :3], 64)), oracle.fps(h[:, :, :3].float().cpu(), 64))
    start = torch.tensor([5, 0, 2047, 17])
    assert np.array_equal(npy(fps_plain(target.cuda(), 32, start_idx=start)), oracle.fps(target[:, :, :3], 32, start=start.numpy()))
    z = torch.cat([torch.zeros(2, 5, 3), torch.rand(2, 300, 3, generator=torch.Generator().manual_seed(2)) + 0.1], dim=1)
    assert np.array_equal(npy(pcl.farthest_point_sample(z.cuda(), 16, skip_origin=True)), oracle.fps(z, 16, skip_origin=True))
    d = torch.tensor([[[0.1, 0.2, 0.3]] * 4 + [[0.9, 0.9, 0.9]]])
    assert npy(fps_plain(d.cuda(), 4))[0].tolist() == [0, 4, 0, 0]
    pts, idx = pcl.sample_farthest_points(target.cuda(), K=50)                                           # utils.py:90 call surface
    assert pts.shape == (4, 50, 6) and torch.equal(pts, torch.gather(target.cuda(), 1, idx.unsqueeze(-1).expand(-1, -1, 6)))


def test_fps_properties_full_size():
    x, _ = synth.uniform_clouds(32, 2048, seed=9)
    idx = fps_plain(x.cuda(), 512)
    assert torch.equal(idx, fps_plain(x.cuda(), 512))                                    # deterministic
    assert (idx[:, 0] == 0).all() and all(len(r.unique()) == 512 for r in idx)
    # prefix property: the first m samples of a longer run are the shorter run
    assert torch.equal(idx[:, :128], fps_plain(x.cuda(), 128))
    # farthest-point samples are spread out: their minimum pairwise distance beats a random subset's
    s = torch.gather(x.cuda(), 1, idx.unsqueeze(-1).expand(-1, -1, 3))
    r = x.cuda()[:, :512]
    mind = lambda p: (torch.cdist(p, p) + 10 * torch.eye(512, device="cuda")).amin(dim=(1, 2))
    assert (mind(s) > 2 * mind(r)).all()


def test_ball_query_vs_oracle_and_reference_golden(golden):
    xyz, new_xyz = torch.from_numpy(golden["bq_xyz"]).cuda(), torch.from_numpy(golden["bq_new_xyz"]).cuda()
    assert np.array_equal(npy(fps_plain(xyz, 128)), golden["bq_fps_idx"])
    for name in "abc":
        radius, nsample = golden[f"bq_{name}_params"]
        got = npy(pcl.query_ball_point(float(radius), int(nsample), xyz, new_xyz))
        assert np.array_equal(got, oracle.ball_query(float(radius), int(nsample), golden["bq_xyz"], golden["bq_new_xyz"]))  # bit-exact
        amb = golden[f"bq_{name}_ambiguous"]
        assert np.array_equal(got[~amb], golden[f"bq_{name}_idx"][~amb])   # == the reference's own query_ball_point off the sphere surface
    # nothing inside the radius -> N everywhere, like the reference (pointnet2_utils.py:107-112)
    far = torch.full((1, 3, 3), 5.0).cuda()
    assert (pcl.query_ball_point(0.1, 4, xyz[:1], far) == xyz.shape[1]).all()
    x, _ = synth.uniform_clouds(3, 3000, seed=4)
    c = x[:, ::7].contiguous()
    assert np.array_equal(npy(pcl.query_ball_point(0.15, 64, x.cuda(), c.cuda())), oracle.ball_query(0.15, 64, x, c))
