"""Development aid: launch the Chamfer forward a few times (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointcloud_b200 as pcl
from pointcloud_b200 import synth
b, n = 32, 2048
x1, t = synth.table_clouds(b, n, seed=0)
x1, x2 = x1.cuda(), t[:, :, :3].contiguous().cuda()
for _ in range(3):
    r = pcl.chamfer_forward_raw(x1, x2)
torch.cuda.synchronize()
print(float(r["loss_xy"].sum()))
