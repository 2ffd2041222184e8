"""CPU tests of the oracles (test infrastructure): C restatements vs independent numpy/torch
restatements, the properties the reference states (emd_module.py:89-95), and the golden vectors produced
by the real reference Python (tests/golden/make_golden.py)."""
import math

import numpy as np
import pytest
import torch

import oracle
from oracle import loss_oracle
from pointcloud_b200 import synth

F32 = np.float32


def fma32(a, b, c):
    # a*b is exact in float64 for float32 inputs; fsum rounds the exact sum once to float64
    return F32(math.fsum([float(a) * float(b), float(c)]))


def emd_numpy(x1, x2, eps, iters):
    """Independent (slow, per-element) restatement of emd_cuda.cu:95-226 for tiny clouds."""
    n = x1.shape[0]
    asg = np.full(n, -1, np.int64); inv = np.full(n, -1, np.int64)
    price = np.zeros(n, F32); max_inc = np.zeros(n, F32); max_idx = np.zeros(n, np.int64)
    bid = np.zeros(n, np.int64); inc = np.zeros(n, F32)
    eps = F32(eps)
    for t in range(iters):
        last = t == iters - 1
        un = [j for j in range(n) if asg[j] == -1]
        if not un:
            break
        for j in un:
            best, better, bi = F32(-1e9), F32(-1e9), -1
            for k in range(n):
                dx, dy, dz = F32(x2[k, 0] - x1[j, 0]), F32(x2[k, 1] - x1[j, 1]), F32(x2[k, 2] - x1[j, 2])
                s = fma32(dz, dz, fma32(dx, dx, F32(dy * dy)))
                r = F32(np.sqrt(s))
                d = F32(3.0 - float(r) - float(price[k]))
                if d > best:
                    better, best, bi = best, d, k
                elif d > better:
                    better = d
            bid[j] = bi; inc[j] = F32(F32(best - better) + eps)
            if inc[j] > max_inc[bi]:
                max_inc[bi] = inc[j]
        for j in un:
            o = bid[j]
            if float(inc[j]) - 1e-6 <= float(max_inc[o]) <= float(inc[j]) + 1e-6:
                max_idx[o] = j
        for j in un:
            o = bid[j]
            if last or max_idx[o] == j:
                if not last and inv[o] != -1:
                    asg[inv[o]] = -1
                inv[o] = j; asg[j] = o
                price[o] = F32(price[o] + inc[j]); max_inc[o] = F32(-1e9)
    dist = np.zeros(n, F32)
    for j in range(n):
        k = asg[j]
        dx, dy, dz = F32(x1[j, 0] - x2[k, 0]), F32(x1[j, 1] - x2[k, 1]), F32(x1[j, 2] - x2[k, 2])
        dist[j] = fma32(dz, dz, fma32(dx, dx, F32(dy * dy)))
    return dist, asg


@pytest.mark.parametrize("n,iters,seed", [(48, 50, 0), (64, 7, 1), (33, 50, 2)])
def test_emd_c_oracle_matches_numpy_restatement(n, iters, seed):
    x1, x2 = synth.uniform_clouds(1, n, seed=seed)
    r = oracle.emd_forward(x1, x2, 0.005, iters)
    d, a = emd_numpy(x1[0].numpy(), x2[0].numpy(), 0.005, iters)
    assert np.array_equal(r["assignment"][0], a)
    assert np.array_equal(r["dist"][0], d)


def test_emd_oracle_properties_reference_demo():
    """emd_module.py:89-95: dist == |x1 - x2[assignment]|^2 ("Verified EMD") and near-bijection."""
    x1, x2 = synth.uniform_clouds(3, 1024, seed=5)
    r = oracle.emd_forward(x1, x2, 0.005, 50, nthreads=3)
    a = r["assignment"].astype(np.int64)
    assert a.min() >= 0 and a.max() < 1024
    matched = np.take_along_axis(x2.numpy(), a[..., None], 1)
    d = ((x1.numpy() - matched) ** 2).sum(-1)
    np.testing.assert_allclose(r["dist"], d, rtol=1e-5, atol=1e-9)
    for i in range(3):
        assert len(np.unique(a[i])) >= 0.93 * 1024  # eps=0.005 / 50 iterations leaves a few duplicates
    # threads only partition the batch
    r1 = oracle.emd_forward(x1, x2, 0.005, 50, nthreads=1)
    assert np.array_equal(r1["assignment"], r["assignment"]) and np.array_equal(r1["dist"], r["dist"])


def test_emd_oracle_converges_to_bijection_with_test_settings():
    """cfg.py:40-41 test setting (eps 0.002, up to 10000 iterations): terminates early on a bijection."""
    x1, x2 = synth.uniform_clouds(1, 256, seed=6)
    r = oracle.emd_forward(x1, x2, 0.002, 10000)
    assert r["iters_run"][0] < 10000
    assert len(np.unique(r["assignment"][0])) == 256


def test_emd_oracle_strided_views_and_backward():
    pred, target = synth.autoencoder_batch(2, 256, seed=7)
    r_view = oracle.emd_forward(pred[:, :, :3], target[:, :, :3], 0.005, 50)
    r_copy = oracle.emd_forward(pred[:, :, :3].contiguous(), target[:, :, :3].contiguous(), 0.005, 50)
    assert np.array_equal(r_view["assignment"], r_copy["assignment"])
    gd = torch.rand(2, 256)
    g1, g2 = oracle.emd_backward(pred[:, :, :3], target[:, :, :3], r_view["assignment"], gd)
    matched = np.take_along_axis(target[:, :, :3].numpy(), r_view["assignment"].astype(np.int64)[..., None], 1)
    np.testing.assert_allclose(g1, 2 * gd.numpy()[..., None] * (pred[:, :, :3].numpy() - matched), rtol=1e-6, atol=1e-9)
    assert not g2.any()  # emd_module.py:69,72


def _torch_chamfer(x, y, xl=None, yl=None):
    n, p1, _ = x.shape; p2 = y.shape[1]
    xl = torch.full((n,), p1) if xl is None else xl; yl = torch.full((n,), p2) if yl is None else yl
    d = ((x[:, :, None, :] - y[:, None, :, :]) ** 2).sum(-1)
    xm = torch.arange(p1)[None] >= xl[:, None]; ym = torch.arange(p2)[None] >= yl[:, None]
    inf = torch.tensor(float("inf"))
    dx = torch.where(ym[:, None, :], inf, d).min(2); dy = torch.where(xm[:, :, None], inf, d).min(1)
    cx = torch.where(xm | (yl[:, None] == 0), torch.zeros(()), dx.values); cy = torch.where(ym | (xl[:, None] == 0), torch.zeros(()), dy.values)
    loss = (cx.sum(1) / xl.clamp(min=1)).sum() / n + (cy.sum(1) / yl.clamp(min=1)).sum() / n
    return loss, cx, cy, dx.indices, dy.indices


@pytest.mark.parametrize("d", [3, 6])
def test_chamfer_oracle_matches_torch_bruteforce(d):
    g = torch.Generator().manual_seed(3)
    x, y = torch.rand(3, 200, d, generator=g), torch.rand(3, 331, d, generator=g)
    yl = torch.tensor([331, 17, 0])
    o = oracle.chamfer_forward(x, y, y_lengths=yl, mode=0)
    loss, cx, cy, ix, iy = _torch_chamfer(x, y, None, yl)
    if d == 3:  # torch sums the 3 squares in the same order -> bit-exact; wider rows are vectorised differently
        assert np.array_equal(o["dist_x"], cx.numpy()) and np.array_equal(o["dist_y"], cy.numpy())
    else:
        np.testing.assert_allclose(o["dist_x"], cx.numpy(), rtol=1e-6); np.testing.assert_allclose(o["dist_y"], cy.numpy(), rtol=1e-6)
    valid_x = (yl > 0)[:, None].expand(-1, 200).numpy()
    assert np.array_equal(o["idx_x"][valid_x], ix.numpy()[valid_x])
    valid_y = (torch.arange(331)[None] < yl[:, None]).numpy()
    assert np.array_equal(o["idx_y"][valid_y], iy.numpy()[valid_y])
    assert abs(float(o["loss"]) - float(loss)) <= 1e-6 * float(loss)
    # padded rows and empty targets stay 0 (SURVEY App. B "Edge")
    assert not o["dist_y"][~valid_y].any() and not o["dist_x"][2].any()


def test_chamfer_oracle_backward_matches_autograd():
    g = torch.Generator().manual_seed(4)
    x, y = torch.rand(2, 150, 3, generator=g).requires_grad_(), torch.rand(2, 180, 3, generator=g).requires_grad_()
    yl = torch.tensor([180, 90])
    loss, *_ = _torch_chamfer(x, y, None, yl)
    (loss * 0.7).backward()
    o = oracle.chamfer_forward(x, y, y_lengths=yl)
    gx, gy = oracle.chamfer_backward(x, y, o["idx_x"], o["idx_y"], 0.7, None, yl)
    np.testing.assert_allclose(gx, x.grad.numpy(), rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(gy, y.grad.numpy(), rtol=2e-5, atol=1e-9)


def test_chamfer_oracle_tie_break_lowest_index_and_modes():
    x = torch.tensor([[[0.5, 0.5, 0.5]]])
    y = torch.tensor([[[0.75, 0.5, 0.5], [0.25, 0.5, 0.5], [0.5, 0.75, 0.5]]])  # three exact ties
    o = oracle.chamfer_forward(x, y)
    assert o["idx_x"][0, 0] == 0
    g = torch.Generator().manual_seed(8)
    a, b = torch.rand(1, 512, 3, generator=g), torch.rand(1, 512, 3, generator=g)
    o0, o1 = oracle.chamfer_forward(a, b, mode=0), oracle.chamfer_forward(a, b, mode=1)
    np.testing.assert_allclose(o0["dist_x"], o1["dist_x"], rtol=1e-6)
    assert (o0["dist_x"] != o1["dist_x"]).any()  # the two roundings really differ
    with pytest.raises(ValueError):
        oracle.chamfer_forward(a, b, y_lengths=torch.tensor([513]))


# ---- golden vectors produced by the real reference Python (tests/golden/make_golden.py) -----------------
def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_loss_oracle_matches_reference_python_autoencoder(golden):
    pred = _t(golden["ae_pred"]).requires_grad_()
    fn = loss_oracle.EarthMoverDistance(eps=0.005, its=50, num_classes=None)
    loss = fn(pred, _t(golden["ae_target"]))
    loss.backward()
    assert float(loss) == pytest.approx(float(golden["ae_loss"]), rel=1e-6)
    assert fn.logged["train_loss/EMD"] == pytest.approx(float(golden["ae_log_EMD"]), rel=1e-6)
    assert fn.logged["train_loss/feature"] == pytest.approx(float(golden["ae_log_feature"]), rel=1e-6)
    np.testing.assert_allclose(pred.grad.numpy(), golden["ae_grad"], rtol=1e-5, atol=1e-10)


def test_loss_oracle_matches_reference_python_segmenter(golden):
    pred = _t(golden["seg_pred"]).requires_grad_()
    fn = loss_oracle.EarthMoverDistance(eps=0.005, its=50, num_classes=5)
    loss = fn(pred, _t(golden["seg_target"]))
    loss.backward()
    assert float(loss) == pytest.approx(float(golden["seg_loss"]), rel=1e-6)
    for k in ("EMD", "feature", "cross_entropy", "kl_divergence"):
        assert fn.logged[f"train_loss/{k}"] == pytest.approx(float(golden[f"seg_log_{k}"]), rel=1e-5)
    np.testing.assert_allclose(pred.grad.numpy(), golden["seg_grad"], rtol=1e-5, atol=1e-10)


def test_emd_oracle_matches_reference_module_golden(golden):
    r = oracle.emd_forward(golden["raw_xyz1"], golden["raw_xyz2"], 0.005, 50)
    assert np.array_equal(r["assignment"], golden["raw_assignment"])
    assert np.array_equal(r["dist"], golden["raw_dist"])


def test_loss_oracle_matches_reference_python_chamfer_family(golden):
    labels = {nm: i for i, nm in enumerate(["env", "cube", "arm", "base", "gripper"])}
    pred = {k: _t(golden[f"mseg_pred_{k}"]).requires_grad_() for k in labels}
    loss = loss_oracle.SegmentingChamferDistance(labels)(pred, _t(golden["mseg_target"]))
    loss.backward()
    assert float(loss) == pytest.approx(float(golden["mseg_loss"]), rel=2e-6)
    for k in labels:
        np.testing.assert_allclose(pred[k].grad.numpy(), golden[f"mseg_grad_{k}"], rtol=2e-5, atol=1e-9)
    p6 = _t(golden["ch6_pred"]).requires_grad_()
    l6 = loss_oracle.ChamferDistance()(p6, _t(golden["ch6_target"]))
    l6.backward()
    assert float(l6) == pytest.approx(float(golden["ch6_loss"]), rel=2e-6)
    np.testing.assert_allclose(p6.grad.numpy(), golden["ch6_grad"], rtol=2e-5, atol=1e-9)
