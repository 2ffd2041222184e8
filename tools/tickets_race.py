"""Development aid: run the ticket path repeatedly on the bench batch and report any deviation from the cluster kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointcloud_b200 as pcl
from pointcloud_b200 import synth
b, n = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for kind in ("table", "noisy"):
    x1, t = synth.table_clouds(b, n, seed=0, regime="independent" if kind == "table" else "noisy")
    x1, x2 = x1.cuda(), t[:, :, :3].contiguous().cuda()
    pcl.set_emd_path("cluster")
    d0, a0, s0 = pcl.emd_forward_raw(x1, x2, 0.005, 50, want_stats=True)
    pcl.set_emd_path("tickets")
    bad = 0
    for r in range(reps):
        d, a, s = pcl.emd_forward_raw(x1, x2, 0.005, 50, want_stats=True)
        torch.cuda.synchronize()
        da, dd = (a != a0).any(1), (d != d0).any(1)
        ds = [(s[:, k] != s0[:, k]).nonzero().flatten().tolist() for k in range(3)]
        if da.any() or dd.any() or any(ds):
            bad += 1
            print(f"{kind} rep {r}: assignment differs in clouds {da.nonzero().flatten().tolist()}, dist in {dd.nonzero().flatten().tolist()}, stats cols {ds}", flush=True)
    print(f"{kind}: {bad} of {reps} runs deviate", flush=True)
