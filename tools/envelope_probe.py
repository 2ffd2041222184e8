"""Development aid: run-to-run / order-to-order spread of the unmodified reference EMD extension vs our kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import oracle
from oracle import build_ref
import pointcloud_b200 as pcl
from pointcloud_b200 import synth
from helpers import ref_emd_forward, npy
ref = build_ref.load_ref()
b, n = 32, 2048
np.set_printoptions(linewidth=250, precision=3)
for regime in ("independent", "noisy"):
    x1, t = synth.table_clouds(b, n, seed=0, regime=regime)
    x2 = t[:, :, :3].contiguous()
    o = oracle.emd_forward(x1, x2, 0.005, 50, nthreads=16)
    d, a, _ = pcl.emd_forward_raw(x1.cuda(), x2.cuda(), 0.005, 50)
    ours = np.sqrt(npy(d).astype(np.float64)).mean(1)
    g = torch.Generator().manual_seed(5)
    runs = []
    for r in range(12):
        perm = torch.arange(n) if r < 4 else torch.randperm(n, generator=g)
        rd, ra = ref_emd_forward(ref, x1[:, perm], x2, 0.005, 50)
        inv = torch.empty_like(perm); inv[perm] = torch.arange(n)
        runs.append(np.sqrt(npy(rd)[:, inv.numpy()].astype(np.float64)).mean(1))
    runs = np.stack(runs)
    # our kernel under the same permutations (its tie rule is index based, so it moves too)
    mine = []
    g = torch.Generator().manual_seed(5)
    for r in range(12):
        perm = torch.arange(n) if r < 4 else torch.randperm(n, generator=g)
        dd, _, _ = pcl.emd_forward_raw(x1[:, perm].cuda(), x2.cuda(), 0.005, 50)
        mine.append(np.sqrt(npy(dd).astype(np.float64)).mean(1))
    mine = np.stack(mine)
    print(f"== {regime}: race events per cloud {o['race_events'].tolist()}")
    print("repeat spread (4 identical runs) per cloud:", (runs[:4].max(0) - runs[:4].min(0)))
    print("order spread (8 permutations)   per cloud:", (runs[4:].max(0) - runs[4:].min(0)))
    print("ours order spread               per cloud:", (mine[4:].max(0) - mine[4:].min(0)))
    print("ours - ref mean                          :", ours - runs.mean(0))
    print("ref std over 12                          :", runs.std(0))
    print("batch: ours", ours.mean(), "ref runs", runs.mean(1), "ours permuted", mine.mean(1))
