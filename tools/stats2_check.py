import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import pointcloud_b200 as pcl
from pointcloud_b200 import synth
import oracle
b, n = 32, 2048
x1, t = synth.table_clouds(b, n, seed=0, regime="independent")
x2 = t[:, :, :3].contiguous()
o = oracle.emd_forward(x1, x2, 0.005, 50, nthreads=16)
pcl.set_emd_path(sys.argv[1])
d, a, s = pcl.emd_forward_raw(x1.cuda(), x2.cuda(), 0.005, 50, want_stats=True)
s = s.cpu().numpy()
print(sys.argv[1], os.environ.get("PCL_EMD_THREADS"), "cs", s[0, 3], "stats2", s[:, 2].tolist())
print("oracle race_events", o["race_events"].tolist())
print("assign ok", np.array_equal(a.cpu().numpy(), o["assignment"]))
