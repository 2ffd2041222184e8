import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pointcloud_b200 as pcl
from pointcloud_b200 import synth
for n in (2048, 8192, 16384):
    x1, t = synth.table_clouds(32, n, seed=0); x2 = t[:, :, :3].contiguous()
    for _ in range(2):
        r = pcl.chamfer_forward_raw(x1.cuda(), x2.cuda())
    torch.cuda.synchronize()
print("ok")
