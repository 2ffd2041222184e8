"""Generate tests/golden/loss_golden.npz by running the REAL reference Python loss code.

Runs only in the build container (needs /root/reference); the GPU box uses the committed .npz.

What is real and what is stubbed
  * REAL: pointcloud_vision/utils.py (EarthMoverDistance, FilteringChamferDistance,
    SegmentingChamferDistance, ChamferDistance, FilterClasses) and
    pointcloud_vision/loss/emd/emd_module.py (emdFunction / emdModule), imported from /root/reference
    unmodified -- i.e. the permutation by assignment, the class histogram / weights, the weighted
    cross-entropy, the MSE feature term, the sqrt-weighted mean, the per-class filter + pad + y_lengths.
  * STUBBED native modules (absent here / need a GPU):
      - `emd` (the pybind module of loss/emd/emd.cpp) -> oracle/emd_oracle.c through the same
        forward/backward signature; the oracle itself is pinned against the unmodified reference
        extension on the GPU box (tests/test_emd_gpu.py);
      - `pytorch3d` -> chamfer_distance restated in plain torch ops (brute force, autograd) -- pytorch3d
        is not installed, so the Chamfer arithmetic stays "parity unpinned" (oracle/chamfer_oracle.c header);
      - `Tensor.cuda()` / `device='cuda'` are redirected to the CPU.
Usage: python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"

import oracle  # noqa: E402
from pointcloud_b200 import synth  # noqa: E402  (input generators only)


def install_stubs():
    # --- emd: same signature as emd.cpp:14-23, backed by the CPU oracle --------------------------------
    emd = types.ModuleType("emd")

    def forward(xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments,
                unass_idx, unass_cnt, unass_cnt_sum, cnt_tmp, max_idx, eps, iters):
        r = oracle.emd_forward(xyz1, xyz2, eps, iters, nthreads=8)
        dist.copy_(torch.from_numpy(r["dist"]))
        assignment.copy_(torch.from_numpy(r["assignment"]))
        return 1

    def backward(xyz1, xyz2, gradxyz, graddist, idx):
        g1, _ = oracle.emd_backward(xyz1, xyz2, idx.numpy(), graddist)
        gradxyz.add_(torch.from_numpy(g1))  # the reference kernel atomically adds into a zeroed buffer
        return 1

    emd.forward, emd.backward = forward, backward
    sys.modules["emd"] = emd

    # --- pytorch3d: only what utils.py imports (utils.py:10-11) --------------------------------------
    p3d = types.ModuleType("pytorch3d")
    ops = types.ModuleType("pytorch3d.ops")
    ops.sample_farthest_points = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    loss = types.ModuleType("pytorch3d.loss")

    def chamfer_distance(x, y, x_lengths=None, y_lengths=None):
        n, p1, _ = x.shape
        p2 = y.shape[1]
        xl = torch.full((n,), p1, dtype=torch.int64) if x_lengths is None else x_lengths.long()
        yl = torch.full((n,), p2, dtype=torch.int64) if y_lengths is None else y_lengths.long()
        d = ((x[:, :, None, :] - y[:, None, :, :]) ** 2).sum(-1)                 # (n, p1, p2) squared L2
        xm = torch.arange(p1)[None] >= xl[:, None]                                # padded rows
        ym = torch.arange(p2)[None] >= yl[:, None]
        inf = torch.tensor(float("inf"))
        cham_x = torch.where(ym[:, None, :], inf, d).min(2).values               # nn of every x in valid y
        cham_y = torch.where(xm[:, :, None], inf, d).min(1).values
        cham_x = torch.where(xm | (yl[:, None] == 0), torch.zeros(()), cham_x)
        cham_y = torch.where(ym | (xl[:, None] == 0), torch.zeros(()), cham_y)
        cham_x = cham_x.sum(1) / xl.clamp(min=1)
        cham_y = cham_y.sum(1) / yl.clamp(min=1)
        return cham_x.sum() / max(n, 1) + cham_y.sum() / max(n, 1), None

    loss.chamfer_distance = chamfer_distance
    p3d.ops, p3d.loss = ops, loss
    sys.modules.update({"pytorch3d": p3d, "pytorch3d.ops": ops, "pytorch3d.loss": loss})

    # --- the package shell: import submodules of /root/reference/pointcloud_vision without running its
    #     __init__.py (which registers gym environments and needs robosuite) -------------------------
    pkg = types.ModuleType("pointcloud_vision")
    pkg.__path__ = [os.path.join(REF, "pointcloud_vision")]
    sys.modules["pointcloud_vision"] = pkg

    # --- CPU redirection of the hard-coded CUDA placement (emd_module.py:43-56,68-69) ----------------
    torch.Tensor.cuda = lambda self, *a, **k: self
    for name in ("zeros",):
        orig = getattr(torch, name)

        def patched(*a, _orig=orig, **k):
            if k.get("device") == "cuda":
                k["device"] = "cpu"
            return _orig(*a, **k)

        setattr(torch, name, patched)


def main():
    install_stubs()
    import pointcloud_vision.utils as ref_utils  # the real reference module

    out = {}
    logs = {}

    def run_emd(tag, pred, target, num_classes):
        pred = pred.clone().requires_grad_()
        loss_fn = ref_utils.EarthMoverDistance(eps=0.005, its=50, num_classes=num_classes)
        logged = {}
        loss_fn.log = lambda k, v: logged.__setitem__(k, float(v))
        loss = loss_fn(pred, target)
        loss.backward()
        out[f"{tag}_pred"], out[f"{tag}_target"] = pred.detach().numpy(), target.numpy()
        out[f"{tag}_loss"] = np.float32(loss.item())
        out[f"{tag}_grad"] = pred.grad.numpy()
        for k, v in logged.items():
            out[f"{tag}_log_{k.split('/')[-1]}"] = np.float32(v)
        logs[tag] = logged

    # (1) Autoencoder loss: EMD(xyz) + MSE(rgb)  (train.py:80-84)
    pred, target = synth.autoencoder_batch(2, 1024, seed=11, regime="independent")
    run_emd("ae", pred, target, None)
    # (2) Segmenter loss: class-weighted EMD + 0.1 * weighted CE  (train.py:98-102)
    pred, target = synth.segmenter_batch(2, 1024, seed=12, regime="noisy")
    run_emd("seg", pred, target, 5)

    # (3) raw emdModule through the reference's autograd Function (emd_module.py:31-79)
    from pointcloud_vision.loss.emd.emd_module import emdModule
    x1, x2 = synth.uniform_clouds(2, 1024, seed=13)
    x1 = x1.requires_grad_()
    dist, asg = emdModule()(x1, x2, 0.005, 50)
    dist.sqrt().mean().backward()
    out.update(raw_xyz1=x1.detach().numpy(), raw_xyz2=x2.numpy(), raw_dist=dist.detach().numpy(),
               raw_assignment=asg.numpy(), raw_grad=x1.grad.numpy())

    # (4) MultiSegmenter loss: SegmentingChamferDistance (train.py:125)
    pred_d, target, labels = synth.multisegmenter_batch(2, 1024, seed=14)
    pred_d = {k: v.clone().requires_grad_() for k, v in pred_d.items()}
    loss = ref_utils.SegmentingChamferDistance(labels)(pred_d, target)
    loss.backward()
    out["mseg_target"] = target.numpy()
    out["mseg_loss"] = np.float32(loss.item())
    for k, v in pred_d.items():
        out[f"mseg_pred_{k}"] = v.detach().numpy()
        out[f"mseg_grad_{k}"] = v.grad.numpy()

    # (5) plain ChamferDistance over all 6 channels (utils.py:209-211)
    g = torch.Generator().manual_seed(15)
    p6, t6 = torch.rand(2, 256, 6, generator=g).requires_grad_(), torch.rand(2, 300, 6, generator=g)
    loss = ref_utils.ChamferDistance()(p6, t6)
    loss.backward()
    out.update(ch6_pred=p6.detach().numpy(), ch6_target=t6.numpy(), ch6_loss=np.float32(loss.item()), ch6_grad=p6.grad.numpy())

    # (6) ball query + sample_and_group through the REAL reference torch code (models/pointnet2_utils.py:93-144);
    #     its FPS is third-party CUDA (pointnet2_ops), stubbed with the CPU oracle of the commented torch algorithm.
    p2o = types.ModuleType("pointnet2_ops")
    p2u = types.ModuleType("pointnet2_ops.pointnet2_utils")
    p2u.furthest_point_sample = lambda xyz, npoint: torch.from_numpy(oracle.fps(xyz, npoint))
    p2o.pointnet2_utils = p2u
    sys.modules.update({"pointnet2_ops": p2o, "pointnet2_ops.pointnet2_utils": p2u})
    import pointcloud_vision.models.pointnet2_utils as ref_p2
    _, t4 = synth.table_clouds(2, 1024, seed=16)
    xyz = t4[:, :, :3].contiguous()
    fps_idx = ref_p2.farthest_point_sample(xyz, 128)
    new_xyz = ref_p2.index_points(xyz, fps_idx)
    out.update(bq_xyz=xyz.numpy(), bq_fps_idx=fps_idx.numpy().astype(np.int32), bq_new_xyz=new_xyz.numpy())
    sq = ref_p2.square_distance(new_xyz, xyz)  # the reference's matmul form: |x|^2 + |y|^2 - 2 x.y, abs error ~3e-7 here
    for name, radius, nsample in (("a", 0.1, 16), ("b", 0.2, 32), ("c", 0.02, 8)):
        gi = ref_p2.query_ball_point(radius, nsample, xyz, new_xyz)
        out[f"bq_{name}_idx"] = gi.numpy().astype(np.int32)
        out[f"bq_{name}_params"] = np.array([radius, nsample], np.float64)
        # centroids with a point so close to the sphere that the matmul form and the difference form may disagree
        out[f"bq_{name}_ambiguous"] = ((sq - radius ** 2).abs() <= 2e-6).any(-1).numpy()

    path = os.path.join(ROOT, "tests", "golden", "loss_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k, v in logs.items():
        print(k, v)
    print({k: float(v) for k, v in out.items() if k.endswith("_loss")})


if __name__ == "__main__":
    main()
