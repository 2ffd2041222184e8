"""Development aid: launch the auction kernel a few times on one input kind (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointcloud_b200 as pcl
from pointcloud_b200 import synth
kind = sys.argv[1] if len(sys.argv) > 1 else "uniform"
b, n = 32, 2048
if kind == "uniform":
    x1, x2 = synth.uniform_clouds(b, n, seed=0)
else:
    x1, t = synth.table_clouds(b, n, seed=0, regime="independent" if kind == "table" else "noisy")
    x2 = t[:, :, :3].contiguous()
x1, x2 = x1.cuda(), x2.cuda()
for _ in range(3):
    d, a, st = pcl.emd_forward_raw(x1, x2, 0.005, 50, want_stats=True)
torch.cuda.synchronize()
print(kind, float(d.sqrt().mean()), st[:, 0].sum().item())
