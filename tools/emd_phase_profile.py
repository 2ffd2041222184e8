"""Development aid: per-phase clock totals of the auction kernel (PCL_EMD_PROFILE=1)."""
import os, sys
os.environ["PCL_EMD_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointcloud_b200 as pcl
from pointcloud_b200 import _lib, synth

L = _lib.lib()
names = ["init", "compact", "lpb_work", "cluster_wait", "resolve", "epilogue", "lpb_seeds", "publish", "wpb_work", "lpb_sync_merge"]
for kind in ("uniform", "table", "noisy"):
    b, n = 32, 2048
    if kind == "uniform":
        x1, x2 = synth.uniform_clouds(b, n, seed=0)
    else:
        x1, t = synth.table_clouds(b, n, seed=0, regime="independent" if kind == "table" else "noisy")
        x2 = t[:, :, :3].contiguous()
    x1, x2 = x1.cuda(), x2.cuda()
    dist = torch.empty(b, n, device="cuda"); asg = torch.empty(b, n, device="cuda", dtype=torch.int32)
    stats = torch.empty(b, 8, device="cuda", dtype=torch.int32)
    wsb = L.pcl_emd_workspace_bytes(b, n); ws = torch.zeros(wsb, device="cuda", dtype=torch.uint8)
    for _ in range(3):
        ws.zero_()
        rc = L.pcl_emd_fwd(*_lib.pts_args(x1), *_lib.pts_args(x2), b, n, 0.005, 50, dist.data_ptr(), asg.data_ptr(), stats.data_ptr(), ws.data_ptr(), wsb, None)
        assert rc == 0
    torch.cuda.synchronize()
    cs = int(stats[0, 3])
    ws = ws[256 + (max(b, 256) * 8 + 255) // 256 * 256:]  # behind the fused-epilogue header of the workspace (csrc/pcl_emd.cu emd_fused_bytes)
    prof = ws.view(torch.int64)[: b * cs * 16].view(b, cs, 16).double().cpu()
    tot = prof.sum(-1)
    print(f"{kind}: cs={cs} iters_run={stats[:,1].tolist()[:6]}.. per-CTA total cycles mean={tot.mean():.0f} max={tot.max():.0f}")
    slow = int(tot.sum(1).argmax())
    print(f"   per-cloud totals (Mcyc): {[round(float(v) / 1e6, 2) for v in tot.mean(1)]}  slowest cloud {slow}: " + ", ".join(f"{nm}={prof[slow,:,i].mean()/1e3:.0f}k" for i, nm in enumerate(names)))
    for i, nm in enumerate(names):
        print(f"   {nm:13s} mean {prof[:,:,i].mean():10.0f} cyc ({100*prof[:,:,i].mean()/tot.mean():5.1f}%)  max {prof[:,:,i].max():10.0f}")
    ev = (stats[:, 4].long() & 0xffffffff) + (stats[:, 5].long() << 32)
    print(f"   executed evals / algorithmic evals = {ev.sum().item() / (stats[:,0].long().sum().item() * n):.3f}")
    it = ws.view(torch.int64)[b * cs * 16: b * cs * 16 + 200].view(50, 4).cpu().tolist()
    print("   per-iteration (U, bid cyc, KS*1000+Gn, wait cyc) of CTA0:", [tuple(int(v) for v in r) for r in it][:50])
    hb = ws.view(torch.int64)[b * cs * 16 + 256: b * cs * 16 + 256 + 200].view(4, 50).cpu()
    rows = []
    for t in range(0, 50):
        best = max(int(hb[c, t]) for c in range(4))
        if best:
            rows.append((t, best >> 40, (best >> 32) & 0xff, (best >> 24) & 0xff, (best >> 16) & 0xff, best & 0xffff))
    print("   slowest warp-per-bidder scan per iteration of cloud 0 (t, cycles, candidate tiles, steps, exact rounds, survivors):", rows)
