"""Development aid: phase totals of the cluster kernel with exported iterations (ticket path) + its workers (PCL_EMD_PROFILE=1)."""
import os, sys
os.environ["PCL_EMD_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointcloud_b200 as pcl
from pointcloud_b200 import _lib, synth

L = _lib.lib()
L.pcl_emd_set_path(3)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = 2048
names = ["init", "compact", "lpb_work", "cluster_wait", "resolve", "epilogue", "lpb_seeds", "publish", "wpb_work", "fetch/merge", "export records", "leftover tickets"]
wn = ["idle/claim", "hdr+load", "run", "finish", "#tasks", "#reloads"]
for kind in ("table", "noisy"):
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
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    off = 256 + (max(b, 256) * 8 + 255) // 256 * 256
    w64 = ws[off:].view(torch.int64)
    prof = w64[: b * cs * 16].view(b, cs, 16).double().cpu()
    tot = prof[:, :, :12].sum(-1)
    print(f"== {kind} B={b} cs={cs}: per-CTA total cycles mean {tot.mean():.0f} max {tot.max():.0f}")
    print("   per-cloud totals (Mcyc):", [round(float(v) / 1e6, 2) for v in tot.mean(1)])
    for i, nm in enumerate(names):
        print(f"   {nm:18s} mean {prof[:, :, i].mean():10.0f} ({100 * prof[:, :, i].mean() / tot.mean():5.1f}%) max {prof[:, :, i].max():10.0f}")
    nw = sm - b * cs
    if nw > 0:
        wk = w64[b * cs * 16 + 512: b * cs * 16 + 512 + nw * 16].view(nw, 16).double().cpu()
        wt = wk[:, :4].sum(1)
        print(f"   workers ({nw}): total cycles mean {wt.mean():.0f}")
        for i, nm in enumerate(wn):
            print(f"   worker {nm:14s} mean {wk[:, i].mean():10.0f} ({100 * wk[:, i].mean() / max(wt.mean(), 1):5.1f}%) max {wk[:, i].max():10.0f}")
