"""Development aid: launch the backward kernels (Chamfer scatter, EMD gather) a few times on the bench workload (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointcloud_b200 as pcl
from pointcloud_b200 import synth
b, n = 32, 2048
x1, t = synth.table_clouds(b, n, seed=0)
x1, x2 = x1.cuda().requires_grad_(), t[:, :, :3].contiguous().cuda()
for _ in range(3):
    loss, _ = pcl.chamfer_distance(x1, x2)
    d, _ = pcl.emdModule()(x1, x2, 0.005, 50)
    (loss + d.sqrt().mean()).backward()
torch.cuda.synchronize()
print(float(loss))
