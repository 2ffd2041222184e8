"""Development aid: cluster / tickets / team at N=2048 for the batch sizes given on the command line (PCL_EMD_CS forces the cluster size)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointcloud_b200 as pcl
from pointcloud_b200 import synth
from tools.path_sweep import t_of  # noqa

n = 2048
for b in [int(v) for v in sys.argv[1:]]:
    for kind in ("table", "noisy"):
        x1, t = synth.table_clouds(b, n, seed=0, regime="independent" if kind == "table" else "noisy")
        x1, x2 = x1.cuda(), t[:, :, :3].contiguous().cuda()
        pcl.set_emd_path("cluster")
        cs = int(pcl.emd_forward_raw(x1, x2, 0.005, 50, want_stats=True)[2][0, 3])
        r = {p: t_of(x1, x2, p) for p in ("cluster", "tickets")}
        print(f"CS={os.environ.get('PCL_EMD_CS', 'auto'):>4s}->{cs:2d} N={n} {kind:6s} B={b:4d}  cluster {r['cluster']:8.1f}  tickets {r['tickets']:8.1f}", flush=True)
