"""Development aid: auction time of the three kernels (cluster / team / tickets) over batch sizes and cloud sizes -> the AUTO rule."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointcloud_b200 as pcl
from pointcloud_b200 import synth


def t_of(x1, x2, path, reps=5):
    pcl.set_emd_path(path)
    for _ in range(2):
        pcl.emd_forward_raw(x1, x2, 0.005, 50)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        pcl.emd_forward_raw(x1, x2, 0.005, 50)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


if __name__ == "__main__":
  for n in (1024, 2048, 3584):
      for kind in ("table", "noisy"):
          for b in (1, 2, 4, 8, 16, 24, 32, 48, 64, 96, 128):
              x1, t = synth.table_clouds(b, n, seed=0, regime="independent" if kind == "table" else "noisy")
              x1, x2 = x1.cuda(), t[:, :, :3].contiguous().cuda()
              r = {p: t_of(x1, x2, p) for p in ("cluster", "team", "tickets")}
              best = min(r, key=r.get)
              print(f"N={n:5d} {kind:6s} B={b:4d}  cluster {r['cluster']:8.1f}  team {r['team']:8.1f}  tickets {r['tickets']:8.1f}  -> {best} (x{r['cluster'] / r[best]:.2f})", flush=True)
