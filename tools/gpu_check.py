"""Quick GPU bring-up check (development tool, not part of the test-suite): parity of the CUDA path
against the CPU oracles and the unmodified reference EMD extension, plus first timings."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from oracle import build_ref  # noqa: E402
import pointcloud_b200 as pcl  # noqa: E402
from pointcloud_b200 import synth  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def check_chamfer():
    for (b, p1, p2, mode) in [(4, 2048, 2048, "unfused"), (4, 2048, 2048, "fma"), (3, 700, 1300, "unfused"), (2, 21, 2048, "fma")]:
        g = torch.Generator().manual_seed(1)
        x, y = torch.rand(b, p1, 3, generator=g), torch.rand(b, p2, 3, generator=g)
        yl = torch.randint(0, p2 + 1, (b,), generator=g) if p1 != p2 else None
        o = oracle.chamfer_forward(x, y, y_lengths=yl, mode=0 if mode == "unfused" else 1)
        r = pcl.chamfer_forward_raw(x.cuda(), y.cuda(), y_lengths=yl, mode=mode)
        ok = [np.array_equal(r[k].cpu().numpy(), o[k]) for k in ("dist_x", "idx_x", "dist_y", "idx_y")]
        lx = r["loss_xy"].cpu().numpy()
        print(f"chamfer {b}x{p1}x{p2} {mode}: exact(dist_x,idx_x,dist_y,idx_y)={ok} loss gpu={lx.sum():.8f} oracle={o['loss']:.8f}")
        xg, yg = x.cuda().requires_grad_(), y.cuda().requires_grad_()
        loss, _ = pcl.chamfer_distance(xg, yg, y_lengths=yl, mode=mode)
        loss.backward()
        gx, gy = oracle.chamfer_backward(x, y, o["idx_x"], o["idx_y"], 1.0, y_lengths=yl)
        ex = np.abs(xg.grad.cpu().numpy() - gx).max() / max(np.abs(gx).max(), 1e-30)
        ey = np.abs(yg.grad.cpu().numpy() - gy).max() / max(np.abs(gy).max(), 1e-30)
        print(f"   bwd rel err grad_x={ex:.2e} grad_y={ey:.2e}")


def check_emd(ref):
    cases = [("uniform", 2, 1024), ("uniform", 4, 2048), ("table", 4, 2048), ("noisy", 4, 2048), ("uniform", 1, 3072), ("uniform", 3, 1000)]
    for kind, b, n in cases:
        if kind == "uniform":
            x1, x2 = synth.uniform_clouds(b, n, seed=3)
        else:
            x1, t = synth.table_clouds(b, n, seed=3, regime="independent" if kind == "table" else "noisy")
            x2 = t[:, :, :3].contiguous()
        o = oracle.emd_forward(x1, x2, 0.005, 50, nthreads=8)
        d, a, st = pcl.emd_forward_raw(x1.cuda(), x2.cuda(), 0.005, 50, want_stats=True)
        torch.cuda.synchronize()
        st = st.cpu().numpy()
        ok_d = np.array_equal(d.cpu().numpy(), o["dist"]); ok_a = np.array_equal(a.cpu().numpy(), o["assignment"])
        print(f"emd {kind} {b}x{n}: vs oracle exact dist={ok_d} asg={ok_a} | sumU gpu={st[:,0].tolist()} oracle={o['sum_unass'].tolist()} "
              f"races(oracle objs)={o['race_events'].tolist()} extraq(gpu)={st[:,2].tolist()} cs={st[0,3]}")
        if not ok_a:
            print("    mismatching entries per cloud:", (a.cpu().numpy() != o["assignment"]).sum(1).tolist())
        if ref is not None and n % 1024 == 0:
            rd, ra = ref_emd(ref, x1.cuda(), x2.cuda(), 0.005, 50)
            same = [(bool(np.array_equal(ra[i].cpu().numpy(), o["assignment"][i])), bool(np.array_equal(rd[i].cpu().numpy(), o["dist"][i]))) for i in range(b)]
            print(f"    reference ext vs oracle per cloud (asg,dist): {same}")


def ref_emd(ref, xyz1, xyz2, eps, iters):
    b, n, _ = xyz1.shape
    dev = 'cuda'
    dist = torch.zeros(b, n, device=dev)
    assignment = torch.zeros(b, n, device=dev, dtype=torch.int32) - 1
    assignment_inv = torch.zeros(b, n, device=dev, dtype=torch.int32) - 1
    price = torch.zeros(b, n, device=dev)
    bid = torch.zeros(b, n, device=dev, dtype=torch.int32)
    bid_increments = torch.zeros(b, n, device=dev)
    max_increments = torch.zeros(b, n, device=dev)
    unass_idx = torch.zeros(b * n, device=dev, dtype=torch.int32)
    max_idx = torch.zeros(b * n, device=dev, dtype=torch.int32)
    unass_cnt = torch.zeros(512, dtype=torch.int32, device=dev)
    unass_cnt_sum = torch.zeros(512, dtype=torch.int32, device=dev)
    cnt_tmp = torch.zeros(512, dtype=torch.int32, device=dev)
    ref.forward(xyz1.contiguous(), xyz2.contiguous(), dist, assignment, price, assignment_inv, bid, bid_increments,
                max_increments, unass_idx, unass_cnt, unass_cnt_sum, cnt_tmp, max_idx, eps, iters)
    return dist, assignment


def timings(ref):
    b, n = 32, 2048
    for kind in ("uniform", "table", "noisy"):
        if kind == "uniform":
            x1, x2 = synth.uniform_clouds(b, n, seed=0)
        else:
            x1, t = synth.table_clouds(b, n, seed=0, regime="independent" if kind == "table" else "noisy")
            x2 = t[:, :, :3].contiguous()
        x1, x2 = x1.cuda(), x2.cuda()
        t_emd = timeit(lambda: pcl.emd_forward_raw(x1, x2, 0.005, 50))
        _, _, st = pcl.emd_forward_raw(x1, x2, 0.005, 50, want_stats=True)
        su = st[:, 0].sum().item()
        ev = ((st[:, 4].long() & 0xffffffff) + (st[:, 5].long() << 32)).sum().item()
        msg = f"time {kind} B=32 N=2048: emd_fwd {t_emd*1e3:.1f} us  sumU={su} -> {su*n/t_emd/1e6:.1f} G pair-evals/s (executed {ev/(su*n):.3f})"
        if ref is not None:
            t_ref = timeit(lambda: ref_emd(ref, x1, x2, 0.005, 50), iters=5, warm=1)
            msg += f" | reference ext {t_ref*1e3:.1f} us"
        t_cf = timeit(lambda: pcl.chamfer_forward_raw(x1, x2))
        xg = x1.clone().requires_grad_()
        def fb():
            xg.grad = None
            l, _ = pcl.chamfer_distance(xg, x2)
            l.backward()
        t_cfb = timeit(fb)
        msg += f" | chamfer fwd {t_cf*1e3:.1f} us, fwd+bwd {t_cfb*1e3:.1f} us"
        print(msg)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    ref = build_ref.load_ref()
    print("reference ext:", "loaded" if ref is not None else "absent")
    check_chamfer()
    check_emd(ref)
    timings(ref)
