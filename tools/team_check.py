"""Development aid: team kernel (pcl_emd_team.cu) against the cluster kernel (pcl_emd.cu): outputs must be identical; times of both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointcloud_b200 as pcl
from pointcloud_b200 import _lib, synth

L = _lib.lib()
tag = sys.argv[1] if len(sys.argv) > 1 else ""
alt = int(os.environ.get("ALT_PATH", "3"))  # PCL_EMD_PATH_* compared with the plain cluster kernel
cases = [("table", 32, 2048), ("noisy", 32, 2048), ("uniform", 32, 2048), ("table", 4, 2048), ("noisy", 4, 2048), ("table", 1, 2048),
         ("table", 8, 2048), ("table", 16, 2048), ("table", 64, 2048), ("noisy", 64, 2048), ("table", 100, 2048), ("uniform", 5, 1000),
         ("uniform", 2, 333), ("uniform", 3, 37), ("table", 8, 3584), ("uniform", 40, 1024)]
if len(sys.argv) > 2:
    cases = [c for c in cases if c[1] == int(sys.argv[2])]


def run(x1, x2, path, reps):
    assert L.pcl_emd_set_path(path) == 0
    for _ in range(2):
        d, a, st = pcl.emd_forward_raw(x1, x2, 0.005, 50, want_stats=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        pcl.emd_forward_raw(x1, x2, 0.005, 50)
    e1.record()
    torch.cuda.synchronize()
    return d, a, st, e0.elapsed_time(e1) / reps * 1e3


for kind, b, n in cases:
    if kind == "uniform":
        x1, x2 = synth.uniform_clouds(b, n, seed=0)
    else:
        x1, t = synth.table_clouds(b, n, seed=0, regime="independent" if kind == "table" else "noisy")
        x2 = t[:, :, :3].contiguous()
    x1, x2 = x1.cuda(), x2.cuda()
    dc, ac, sc, tc = run(x1, x2, 1, 10)
    dt, at, stt, tt = run(x1, x2, alt, 10)
    ok = bool((ac == at).all()) and bool((dc == dt).all()) and bool((sc[:, :3] == stt[:, :3]).all())
    ev = lambda s: ((s[:, 4].long() & 0xffffffff) + (s[:, 5].long() << 32)).sum().item()
    print(f"[{tag}] {kind:8s} B={b:3d} N={n:5d} cluster(cs={int(sc[0,3])}) {tc:8.1f} us | path{alt} {tt:8.1f} us | x{tc/tt:5.2f} | identical={ok} | evals {ev(sc):.3e} / {ev(stt):.3e}", flush=True)
L.pcl_emd_set_path(0)
