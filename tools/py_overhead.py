"""Development aid: where does the Python loss path spend host time?  (config 2 shapes, B=32, N=2048)

Prints, per call pattern, the wall time per step with a device synchronise after every step (host + device) and the
pure device time (CUDA events, steps queued back to back), so that host-bound paths show up as wall >> device."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointcloud_b200 as pcl
from pointcloud_b200 import synth

B, N = 32, 2048
dev = torch.device("cuda")


def measure(name, fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
        torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / iters
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / iters
    torch.cuda.synchronize()
    print(f"{name:58s} wall+sync {wall*1e6:8.1f} us | queued: device {e0.elapsed_time(e1)/iters*1e3:8.1f} us, host issue {host*1e6:8.1f} us")


for regime in ("noisy", "independent"):
    p, t = synth.table_clouds(B, N, seed=0, regime=regime)
    p, t3 = p.to(dev), t[:, :, :3].contiguous().to(dev)
    pa, ta = synth.autoencoder_batch(B, N, seed=0, regime=regime)
    pa, ta = pa.to(dev), ta.to(dev)
    ps, ts = synth.segmenter_batch(B, N, seed=0, regime=regime)
    ps, ts = ps.to(dev), ts.to(dev)
    emd = pcl.emdModule()
    print(f"--- regime {regime}")
    measure("emd_forward_raw (C ABI + 3 torch.empty)", lambda: pcl.emd_forward_raw(p, t3, 0.005, 50))
    measure("chamfer_forward_raw", lambda: pcl.chamfer_forward_raw(p, t3))

    def emd_step():
        x = p.detach().requires_grad_()
        d, _ = emd(x, t3, 0.005, 50)
        d.sqrt().mean().backward()
    measure("emdModule fwd + sqrt.mean + bwd", emd_step)

    def ch_step():
        x = p.detach().requires_grad_()
        pcl.chamfer_distance(x, t3)[0].backward()
    measure("chamfer_distance fwd + bwd", ch_step)

    def both():
        x = p.detach().requires_grad_()
        c, _ = pcl.chamfer_distance(x, t3)
        d, _ = emd(x, t3, 0.005, 50)
        (c + d.sqrt().mean()).backward()
    measure("bench python_api step (chamfer + emd, one backward)", both)

    ae = pcl.EarthMoverDistance(0.005, 50)
    def ae_step():
        x = pa.detach().requires_grad_()
        ae(x, ta).backward()
    measure("EarthMoverDistance (Autoencoder loss) fwd + bwd", ae_step)
    sg = pcl.EarthMoverDistance(0.005, 50, num_classes=5)
    def sg_step():
        x = ps.detach().requires_grad_()
        sg(x, ts).backward()
    measure("EarthMoverDistance (Segmenter loss) fwd + bwd", sg_step)
    sh = pcl.ShardedLoss(pcl.EarthMoverDistance(0.005, 50, num_classes=5))
    def sh_step():
        x = ps.detach().requires_grad_()
        sh(x, ts).backward()
    measure("ShardedLoss(Segmenter loss), 1 rank", sh_step)
