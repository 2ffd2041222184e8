"""Small end-to-end case for compute-sanitizer (memcheck): every kernel of the library once, tiny sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointcloud_b200 as pcl
from pointcloud_b200 import synth
for n in (256, 1000):
    pred, target = synth.segmenter_batch(3, n, seed=1)
    p = pred.cuda().requires_grad_()
    loss = pcl.EarthMoverDistance(0.005, 20, num_classes=5)(p, target.cuda())
    loss.backward()
    pa, ta = synth.autoencoder_batch(2, n, seed=2)
    q = pa.cuda().half().requires_grad_()
    l2 = pcl.EarthMoverDistance(0.005, 20)(q, ta.cuda()) + pcl.ChamferDistance()(q.float(), ta.cuda())
    l2.backward()
    x = torch.rand(2, 333, 3).cuda().requires_grad_()
    l3, _ = pcl.chamfer_distance(x, torch.rand(2, 777, 3).cuda(), y_lengths=torch.tensor([777, 5]).cuda())
    l3.backward()
torch.cuda.synchronize()
print("ok", float(loss), float(l2), float(l3))
