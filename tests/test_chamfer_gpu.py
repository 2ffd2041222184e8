"""GPU parity tests of the Chamfer kernels (through the C ABI) against the CPU oracle (bit-exact distances
and indices, scalar loss / gradients within 1e-5 relative -- BASELINE.json north_star), golden vectors, and
size-independent properties at full size.  Chamfer parity is UNPINNED with respect to pytorch3d itself
(not available); see oracle/chamfer_oracle.c."""
import numpy as np
import pytest
import torch

import oracle
import pointcloud_b200 as pcl
from pointcloud_b200 import synth
from helpers import npy

pytestmark = pytest.mark.gpu
REL = 1e-5  # north star tolerance for fp32 scalars and gradients


def run_case(x, y, xl=None, yl=None, mode="unfused"):
    o = oracle.chamfer_forward(x, y, xl, yl, mode=0 if mode == "unfused" else 1, nthreads=8)
    r = pcl.chamfer_forward_raw(x.cuda(), y.cuda(), xl, yl, mode=mode)
    for k in ("dist_x", "dist_y", "idx_x", "idx_y"):
        assert np.array_equal(npy(r[k]), o[k]), k          # bit-exact fp32 distances, exact indices
    lx = npy(r["loss_xy"])
    assert abs(lx[0] - o["loss_x"]) <= REL * max(o["loss_x"], 1e-30) and abs(lx[1] - o["loss_y"]) <= REL * max(o["loss_y"], 1e-30)
    return o, r


@pytest.mark.parametrize("mode", ["unfused", "fma"])
@pytest.mark.parametrize("b,p1,p2", [(4, 2048, 2048), (3, 700, 1300), (2, 21, 2048), (1, 1, 1), (5, 1024, 33), (2, 4100, 4099)])
def test_chamfer_forward_bit_exact_vs_oracle(b, p1, p2, mode):
    g = torch.Generator().manual_seed(p1 + p2)
    run_case(torch.rand(b, p1, 3, generator=g), torch.rand(b, p2, 3, generator=g), mode=mode)


def test_chamfer_table_shaped_and_ties():
    x, t = synth.table_clouds(4, 2048, seed=2)
    run_case(x, t[:, :, :3].contiguous())
    # duplicated targets => exact ties => lowest index must win (knn: strict '<')
    g = torch.Generator().manual_seed(5)
    y = torch.rand(2, 512, 3, generator=g)
    y2 = torch.cat([y, y], dim=1)
    o, r = run_case(torch.rand(2, 300, 3, generator=g), y2)
    assert (npy(r["idx_x"]) < 512).all()
    # coordinates on a coarse grid produce many equal distances
    xg = torch.randint(0, 8, (2, 600, 3), generator=g).float() / 8
    yg = torch.randint(0, 8, (2, 700, 3), generator=g).float() / 8
    run_case(xg, yg)


def test_chamfer_filter_stress():
    """The kernel scans an approximate expanded-form distance and re-scans candidate chunks exactly; these inputs
    make the approximation useless (far from the origin, huge / mixed scales, all points equal) or non-finite."""
    g = torch.Generator().manual_seed(13)
    x, y = torch.rand(2, 700, 3, generator=g), torch.rand(2, 1100, 3, generator=g)
    for off in (10.0, 1000.0, -3.0e4):          # |t|^2 >> nearest-neighbour distances: every chunk is a candidate
        run_case(x + off, y + off)
        run_case(x * 0.01 + off, y * 0.01 + off, mode="fma")
    run_case(x * 1e-3, y * 1e-3)
    run_case(x * 1e-18, y * 1e-18)                # squared distances underflow towards denormals
    run_case(x * 1e15, y * 1e15)                  # |t|^2 = 1e30: large but finite
    run_case(x * 1e19, y * 1e19)                  # |t|^2 overflows to +inf while the exact distances stay finite
    scale = torch.logspace(-6, 6, 1100).view(1, -1, 1)
    run_case(x, y * scale)                        # one cloud mixes 12 orders of magnitude
    run_case(torch.full((2, 300, 3), 0.25), torch.full((2, 2000, 3), 0.25))  # all equal: index 0 everywhere
    z = y.clone(); z[:, 5::7] = z[:, 4::7][:, : z[:, 5::7].shape[1]]          # many exact duplicates
    run_case(x, z)
    yn = y.clone(); yn[0, 17, 1] = float("nan"); yn[1, 3, 0] = float("inf"); yn[1, 600, 2] = -float("inf")
    o = oracle.chamfer_forward(x, yn, mode=0)
    r = pcl.chamfer_forward_raw(x.cuda(), yn.cuda())
    assert np.array_equal(npy(r["idx_x"]), o["idx_x"]) and np.array_equal(npy(r["dist_x"]), o["dist_x"])   # non-finite targets are never nearest


def test_chamfer_variable_lengths_and_empty():
    g = torch.Generator().manual_seed(7)
    x, y = torch.rand(4, 820, 3, generator=g), torch.rand(4, 1500, 3, generator=g)
    yl = torch.tensor([1500, 0, 1, 777])       # class absent in a cloud => length 0 (utils.py:222-228)
    xl = torch.tensor([820, 820, 5, 0])
    o, r = run_case(x, y, None, yl)
    assert not npy(r["dist_x"])[1].any() and not npy(r["dist_y"])[1].any()
    run_case(x, y, xl, yl)
    run_case(x, y, xl, None, mode="fma")


@pytest.mark.parametrize("d", [1, 2, 4, 6, 8])
def test_chamfer_generic_feature_width(d):
    g = torch.Generator().manual_seed(d)
    run_case(torch.rand(2, 300, d, generator=g), torch.rand(2, 257, d, generator=g))   # ChamferDistance over all channels (utils.py:209-211)


def test_chamfer_backward_vs_oracle_and_autograd_surface():
    g = torch.Generator().manual_seed(11)
    x, y = torch.rand(3, 600, 3, generator=g), torch.rand(3, 900, 3, generator=g)
    yl = torch.tensor([900, 450, 0])
    o = oracle.chamfer_forward(x, y, None, yl)
    gx, gy = oracle.chamfer_backward(x, y, o["idx_x"], o["idx_y"], 0.37, None, yl)
    xg, yg = x.cuda().requires_grad_(), y.cuda().requires_grad_()
    loss, normals = pcl.chamfer_distance(xg, yg, y_lengths=yl.cuda())
    assert normals is None and loss.dim() == 0
    assert abs(float(loss) - float(o["loss"])) <= REL * float(o["loss"])
    (loss * 0.37).backward()
    np.testing.assert_allclose(npy(xg.grad), gx, rtol=REL, atol=1e-9)
    np.testing.assert_allclose(npy(yg.grad), gy, rtol=REL, atol=1e-9)
    assert not npy(yg.grad)[1, 450:].any() and not npy(yg.grad)[2].any()      # padded targets get no gradient


def test_chamfer_strided_and_half_inputs():
    pred, target = synth.autoencoder_batch(2, 1024, seed=3)
    o = oracle.chamfer_forward(pred[:, :, :3], target[:, :, :3])
    r = pcl.chamfer_forward_raw(pred.cuda()[:, :, :3], target.cuda()[:, :, :3])
    assert np.array_equal(npy(r["idx_x"]), o["idx_x"]) and np.array_equal(npy(r["dist_y"]), o["dist_y"])
    ph = pred.cuda().half()
    oh = oracle.chamfer_forward(ph[:, :, :3].float().cpu(), target[:, :, :3])
    rh = pcl.chamfer_forward_raw(ph[:, :, :3], target.cuda()[:, :, :3])
    assert np.array_equal(npy(rh["idx_x"]), oh["idx_x"]) and np.array_equal(npy(rh["dist_x"]), oh["dist_x"])


def test_chamfer_full_size_properties():
    """BASELINE config 2 (B=32, N=M=2048) and the large end of config 5 (N=16384): properties only."""
    for b, n in [(32, 2048), (4, 16384)]:
        x, y = synth.uniform_clouds(b, n, seed=1)
        xc, yc = x.cuda(), y.cuda()
        r = pcl.chamfer_forward_raw(xc, yc)
        # (1) every reported distance is the distance to the reported index, bit-exactly
        m = torch.gather(yc, 1, r["idx_x"].long().unsqueeze(-1).expand(-1, -1, 3))
        df = xc - m
        d = (df[..., 0] * df[..., 0] + df[..., 1] * df[..., 1]) + df[..., 2] * df[..., 2]
        assert torch.equal(d, r["dist_x"])
        # (2) no sampled target is closer than the reported nearest neighbour
        probe = yc[:, torch.randint(0, n, (64,), generator=torch.Generator().manual_seed(2))]
        dp = ((xc[:, :, None, :] - probe[:, None, :, :]) ** 2).sum(-1).min(2).values
        assert (r["dist_x"] <= dp * (1 + 1e-6)).all()
        # (3) swapping the arguments swaps the outputs; (4) chamfer(x, x) == 0 with the identity match
        s = pcl.chamfer_forward_raw(yc, xc)
        assert torch.equal(s["dist_y"], r["dist_x"]) and torch.equal(s["idx_y"], r["idx_x"]) and torch.equal(s["dist_x"], r["dist_y"])
        z = pcl.chamfer_forward_raw(xc, xc)
        assert not z["dist_x"].any() and float(z["loss_xy"].sum()) == 0.0
        # (5) permuting the targets leaves the distances unchanged and maps the indices through the permutation
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(3)).cuda()
        pr = pcl.chamfer_forward_raw(xc, yc[:, perm])
        assert torch.equal(pr["dist_x"], r["dist_x"])
        assert torch.equal(perm[pr["idx_x"].long()], r["idx_x"].long()) or (perm[pr["idx_x"].long()] != r["idx_x"].long()).float().mean() < 1e-4


def test_chamfer_golden_and_argument_checks(golden):
    p6 = torch.from_numpy(golden["ch6_pred"]).cuda().requires_grad_()
    l6 = pcl.ChamferDistance()(p6, torch.from_numpy(golden["ch6_target"]).cuda())
    l6.backward()
    assert float(l6) == pytest.approx(float(golden["ch6_loss"]), rel=REL)
    np.testing.assert_allclose(npy(p6.grad), golden["ch6_grad"], rtol=REL, atol=1e-9)
    x = torch.rand(2, 10, 3).cuda()
    with pytest.raises(ValueError):
        pcl.chamfer_distance(x, torch.rand(3, 10, 3).cuda())
    with pytest.raises(ValueError):
        pcl.chamfer_distance(x, x, y_lengths=torch.tensor([1, 2, 3]).cuda())
    with pytest.raises(NotImplementedError):
        pcl.chamfer_distance(x, x, point_reduction="sum")
    from pointcloud_b200._lib import PclError
    with pytest.raises(PclError):
        pcl.chamfer_distance(torch.rand(1, 4, 9).cuda(), torch.rand(1, 4, 9).cuda())   # D > 8
    l0, _ = pcl.chamfer_distance(torch.rand(0, 5, 3).cuda(), torch.rand(0, 7, 3).cuda())
    assert float(l0) == 0.0


@pytest.fixture
def pruned_everywhere():
    """Forces the spatially pruned nearest-neighbour path (Morton order + tile boxes) for every cloud size, restores the default after."""
    L = pcl._lib.lib()
    L.pcl_chamfer_set_prune_min(1)
    yield
    L.pcl_chamfer_set_prune_min(-1)


def test_pruned_path_is_bit_exact_vs_oracle(pruned_everywhere):
    """Same cases as the brute-force kernel's tests, through the pruned path: random, Table-shaped, exact ties (duplicates, lattice),
    ragged lengths incl. empty clouds, far from the origin / tiny / huge scales, all points equal, non-finite coordinates."""
    g = torch.Generator().manual_seed(101)
    for b, p1, p2 in [(4, 2048, 2048), (3, 700, 1300), (2, 21, 2048), (1, 1, 1), (5, 1024, 33), (2, 4100, 4099)]:
        for mode in ("unfused", "fma"):
            run_case(torch.rand(b, p1, 3, generator=g), torch.rand(b, p2, 3, generator=g), mode=mode)
    x, t = synth.table_clouds(4, 2048, seed=2)
    run_case(x, t[:, :, :3].contiguous())
    y = torch.rand(2, 512, 3, generator=g)
    o, r = run_case(torch.rand(2, 300, 3, generator=g), torch.cat([y, y], dim=1))
    assert (npy(r["idx_x"]) < 512).all()                      # duplicated targets: the lowest index wins
    run_case(torch.randint(0, 8, (2, 600, 3), generator=g).float() / 8, torch.randint(0, 8, (2, 700, 3), generator=g).float() / 8)
    x, y = torch.rand(2, 700, 3, generator=g), torch.rand(2, 1100, 3, generator=g)
    for off in (10.0, 1000.0, -3.0e4):
        run_case(x + off, y + off)
    run_case(x * 1e-18, y * 1e-18)
    run_case(x * 1e15, y * 1e15)
    run_case(torch.full((2, 300, 3), 0.25), torch.full((2, 500, 3), 0.25))      # every point in one cell
    xl, yl = torch.tensor([700, 0]), torch.tensor([5, 1100])
    run_case(x, y, xl, yl)
    xn = x.clone(); xn[0, 3, 1] = float("nan"); xn[1, 10, 0] = float("inf")
    yn = y.clone(); yn[0, 7, 2] = float("-inf"); yn[1, 0, 0] = float("nan")
    o, r = oracle.chamfer_forward(x, yn, mode=0), pcl.chamfer_forward_raw(x.cuda(), yn.cuda())
    assert np.array_equal(npy(r["idx_x"]), o["idx_x"]) and np.array_equal(npy(r["dist_x"]), o["dist_x"])   # non-finite targets are never nearest
    o, r = oracle.chamfer_forward(xn, y, mode=0), pcl.chamfer_forward_raw(xn.cuda(), y.cuda())
    assert np.array_equal(npy(r["idx_y"]), o["idx_y"]) and np.array_equal(npy(r["dist_y"]), o["dist_y"])
    assert np.array_equal(npy(r["idx_x"]), o["idx_x"]) and np.array_equal(npy(r["dist_x"]), o["dist_x"], equal_nan=True)  # non-finite queries


def test_pruned_path_equals_brute_force_at_config5_sizes():
    """BASELINE config 5's large clouds (the sizes that take the pruned path by default): both implementations must return identical
    distances and indices (the CPU oracle needs minutes there; it is checked at N=8192 on two clouds)."""
    L = pcl._lib.lib()
    for n in (8192, 16384):
        xu, yu = synth.uniform_clouds(4, n, seed=n)
        xt, tt = synth.table_clouds(4, n, seed=n + 1)
        for x, y in ((xu, yu), (xt, tt[:, :, :3].contiguous())):
            L.pcl_chamfer_set_prune_min(0)
            a = pcl.chamfer_forward_raw(x.cuda(), y.cuda())
            L.pcl_chamfer_set_prune_min(-1)
            b = pcl.chamfer_forward_raw(x.cuda(), y.cuda())
            for k in ("dist_x", "dist_y", "idx_x", "idx_y"):
                assert torch.equal(a[k], b[k]), (n, k)
            assert torch.allclose(a["loss_xy"], b["loss_xy"], rtol=1e-5)
    x, y = synth.uniform_clouds(2, 8192, seed=3)
    run_case(x, y)
