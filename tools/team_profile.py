"""Development aid: phase totals of the team kernel (PCL_EMD_PROFILE=1): owners and workers."""
import os, sys
os.environ["PCL_EMD_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointcloud_b200 as pcl
from pointcloud_b200 import _lib, synth

L = _lib.lib()
L.pcl_emd_set_path(2)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = 2048
on = ["setup", "compact", "records+publish", "own tasks", "wait workers", "fetch bids", "resolve", "local scan", "epilogue", "#own tasks", "#tasks", "-"]
wn = ["idle/claim", "hdr+load", "run", "finish", "#tasks", "#reloads", "-", "-"]
for kind in ("table", "noisy"):
    x1, t = synth.table_clouds(b, n, seed=0, regime="independent" if kind == "table" else "noisy")
    x2 = t[:, :, :3].contiguous()
    x1, x2 = x1.cuda(), x2.cuda()
    dist = torch.empty(b, n, device="cuda"); asg = torch.empty(b, n, device="cuda", dtype=torch.int32)
    stats = torch.empty(b, 8, device="cuda", dtype=torch.int32)
    wsb = L.pcl_emd_workspace_bytes(b, n); ws = torch.zeros(wsb, device="cuda", dtype=torch.uint8)
    for _ in range(3):
        rc = L.pcl_emd_fwd(*_lib.pts_args(x1), *_lib.pts_args(x2), b, n, 0.005, 50, dist.data_ptr(), asg.data_ptr(), stats.data_ptr(), ws.data_ptr(), wsb, None)
        assert rc == 0, _lib.last_error() if hasattr(_lib, "last_error") else rc
    torch.cuda.synchronize()
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    off = 256 + (max(b, 256) * 8 + 255) // 256 * 256
    prof = ws[off:].view(torch.int64)[: sm * 16].view(sm, 16).double().cpu()
    own, wk = prof[:b, :12], prof[b:, :8]
    tot = own[:, :9].sum(1)
    print(f"== {kind} B={b}: owners total cycles mean {tot.mean():.0f} max {tot.max():.0f} min {tot.min():.0f}")
    for i, nm in enumerate(on[:11]):
        print(f"   owner {nm:16s} mean {own[:, i].mean():10.0f} ({100 * own[:, i].mean() / tot.mean():5.1f}%) max {own[:, i].max():10.0f}")
    if wk.shape[0]:
        wt = wk[:, :4].sum(1)
        print(f"   workers ({wk.shape[0]}): total cycles mean {wt.mean():.0f}")
        for i, nm in enumerate(wn[:6]):
            print(f"   worker {nm:14s} mean {wk[:, i].mean():10.0f} ({100 * wk[:, i].mean() / wt.mean():5.1f}%) max {wk[:, i].max():10.0f}")
