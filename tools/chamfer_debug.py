import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
import pointcloud_b200 as pcl
for (b, p1, p2) in [(1, 64, 32), (1, 128, 64), (1, 128, 512), (1, 128, 1024), (2, 2048, 2048)]:
    g = torch.Generator().manual_seed(p1 + p2)
    x, y = torch.rand(b, p1, 3, generator=g), torch.rand(b, p2, 3, generator=g)
    o = oracle.chamfer_forward(x, y, None, None, mode=0, nthreads=4)
    r = pcl.chamfer_forward_raw(x.cuda(), y.cuda(), None, None, mode="unfused")
    for k in ("dist_x", "idx_x", "dist_y", "idx_y"):
        got = r[k].cpu().numpy(); bad = np.argwhere(got != o[k])
        print(b, p1, p2, k, "mismatches", len(bad), "of", got.size)
        for (n, i) in bad[:6]:
            print("   ", n, i, "got", got[n, i], "want", o[k][n, i], "| got idx", r[k.replace('dist', 'idx')].cpu().numpy()[n, i], "want idx", o[k.replace('dist', 'idx')][n, i])
