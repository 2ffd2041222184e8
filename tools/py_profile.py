"""Development aid: cProfile of the Python loss path (host side) on config-2 shapes."""
import cProfile, pstats, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointcloud_b200 as pcl
from pointcloud_b200 import synth
p, t = synth.table_clouds(32, 2048, seed=0, regime="noisy")
p, t3 = p.cuda(), t[:, :, :3].contiguous().cuda()
emd = pcl.emdModule()
def both():
    x = p.detach().requires_grad_()
    c, _ = pcl.chamfer_distance(x, t3)
    d, _ = emd(x, t3, 0.005, 50)
    (c + d.sqrt().mean()).backward()
for _ in range(10): both()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(200): both()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
