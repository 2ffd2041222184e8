"""Development aid: composite step (pcl_chamfer_emd_step) time per regime + Chamfer forward alone, B=32, N=2048."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pointcloud_b200 as pcl
from pointcloud_b200 import synth
tag = sys.argv[1] if len(sys.argv) > 1 else ""
b, n = 32, 2048
step = pcl.ShardedChamferEmdStep(b, n, torch.device("cuda"), 0.005, 50)
out = []
for regime in ("independent", "noisy"):
    sets = []
    for s in range(8):
        p, t = synth.table_clouds(b, n, seed=s, regime=regime)
        sets.append((p.cuda(), t[:, :, :3].contiguous().cuda()))
    for i in range(5):
        step.step(*sets[i % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40):
        step.step(*sets[i % 8])
    e1.record(); torch.cuda.synchronize()
    out.append(f"{regime} {e0.elapsed_time(e1) / 40 * 1e3:7.1f} us")
p, t = sets[0]
for _ in range(5):
    pcl.chamfer_forward_raw(p, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    pcl.chamfer_forward_raw(p, t)
e1.record(); torch.cuda.synchronize()
out.append(f"chamfer fwd alone {e0.elapsed_time(e1) / 50 * 1e3:6.1f} us")
print(f"[{tag}] " + " | ".join(out), flush=True)
