"""Secondary measurements for the BASELINE.md plan (configs 1, 3, 5; config 2 is bench.py's headline).

  config 1: Chamfer fwd+bwd on the CPU, B=8, N=M=2048 (torch brute force = "reference torch path", and the oracle C port)
  config 3: weighted EMD (Segmenter loss) fwd+bwd, B=32, N=2048, C=5, through the Python loss class
  config 5: Chamfer fwd+bwd, B=64 GLOBAL, N=M in {1024..16384}, on 1/2/4/8 GPUs (C ABI, preallocated outputs; B/G clouds per rank and
            one all-reduce of the two batch sums per step inside the timed region)
Writes one JSON object to stdout.  One GPU: python tools/bench_configs.py > gpurun_out/configs.json
N GPUs (config 5 only): python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_configs.py --only 5
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (CPU baseline leg only)
import pointcloud_b200 as pcl  # noqa: E402
from pointcloud_b200 import _lib, synth  # noqa: E402


def ev_time(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def config1():
    b, n = 8, 2048
    pred, target = synth.table_clouds(b, n, seed=0)
    x, y = pred, target[:, :, :3].contiguous()
    out = {"shape": [b, n, n], "cpu_count": os.cpu_count()}

    def torch_path():
        xx, yy = x.clone().requires_grad_(), y.clone().requires_grad_()
        d = ((xx[:, :, None, :] - yy[:, None, :, :]) ** 2).sum(-1)
        loss = d.min(2).values.mean(1).mean() + d.min(1).values.mean(1).mean()
        loss.backward()

    for th in (os.cpu_count(), 1):
        torch.set_num_threads(th)
        torch_path()
        t0 = time.perf_counter()
        reps = 3 if th > 1 else 1
        for _ in range(reps):
            torch_path()
        dt = (time.perf_counter() - t0) / reps
        out[f"torch_cpu_{th}_threads_ms"] = dt * 1e3
        out[f"torch_cpu_{th}_threads_clouds_per_s"] = b / dt
    torch.set_num_threads(os.cpu_count())

    def port(th):
        c = oracle.chamfer_forward(x, y, nthreads=th)
        oracle.chamfer_backward(x, y, c["idx_x"], c["idx_y"], 1.0)

    for th in (min(os.cpu_count(), b), 1):
        port(th)
        t0 = time.perf_counter()
        for _ in range(3):
            port(th)
        dt = (time.perf_counter() - t0) / 3
        out[f"oracle_port_{th}_threads_ms"] = dt * 1e3
    # the same batch on the GPU through the public API
    xg, yg = x.cuda(), y.cuda()

    def gpu():
        xx = xg.clone().requires_grad_()
        l, _ = pcl.chamfer_distance(xx, yg)
        l.backward()

    out["gpu_api_ms"] = ev_time(gpu, 50)
    return out


def config3():
    b, n, c = 32, 2048, 5
    out = {"shape": [b, n, 3 + c]}
    for regime in ("independent", "noisy"):
        pred, target = synth.segmenter_batch(b, n, seed=0, regime=regime)
        pg, tg = pred.cuda(), target.cuda()
        fn = pcl.EarthMoverDistance(eps=0.005, its=50, num_classes=c)

        def step():
            p = pg.detach().requires_grad_()
            loss = fn(p, tg)
            loss.backward()

        ms = ev_time(step, 20)
        out[f"{regime}_ms"] = ms
        out[f"{regime}_clouds_per_s"] = b / (ms * 1e-3)
    return out


def config5():
    import torch.distributed as dist
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    L = _lib.lib()
    A = _lib.pts_args
    bg = 64
    b = bg // world
    rows = []
    for n in (1024, 2048, 4096, 8192, 16384):
        xg, yg = synth.uniform_clouds(bg, n, seed=0)
        x, y = xg[rank * b:(rank + 1) * b].cuda(), yg[rank * b:(rank + 1) * b].cuda()
        e = lambda *s, dt=torch.float32: torch.empty(*s, device="cuda", dtype=dt)
        dx, dy, ix, iy, lxy = e(b, n), e(b, n), e(b, n, dt=torch.int32), e(b, n, dt=torch.int32), e(4)
        gx, gy, ones = e(b, n, 3), e(b, n, 3), torch.ones(2, device="cuda")
        wsb = L.pcl_chamfer_workspace_bytes(b, n, n)
        ws = torch.empty(wsb, device="cuda", dtype=torch.uint8)
        st = torch.cuda.current_stream().cuda_stream

        def fwd():
            assert L.pcl_chamfer_fwd(*A(x), None, *A(y), None, b, n, n, 3, 0, dx.data_ptr(), ix.data_ptr(), dy.data_ptr(), iy.data_ptr(),
                                     lxy.data_ptr(), ws.data_ptr(), wsb, st) == 0

        def bwd():
            assert L.pcl_chamfer_bwd(*A(x), None, *A(y), None, b, n, n, 3, ix.data_ptr(), iy.data_ptr(), ones.data_ptr(), gx.data_ptr(),
                                     gy.data_ptr(), st) == 0

        def step():  # the sharded Chamfer step: local kernels + the one collective on the two batch sums
            fwd()
            if world > 1:
                dist.all_reduce(lxy[2:4])
            bwd()

        it = 50 if n <= 4096 else 10
        if world > 1:
            dist.barrier()
        tf, tb, ts = ev_time(fwd, it), ev_time(bwd, it), ev_time(step, it)
        if world > 1:
            t = torch.tensor([tf, tb, ts], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tf, tb, ts = (float(v) for v in t)
        evals = 2.0 * b * n * n
        rows.append({"N": n, "clouds_per_gpu": b, "fwd_ms": tf, "bwd_ms": tb, "step_ms": ts, "clouds_per_s": bg / (ts * 1e-3),
                     "directed_pair_evals_per_s_per_gpu": evals / (tf * 1e-3), "fp32_flop_frac_of_74.4T": 8 * evals / (tf * 1e-3) / 74.45e12,
                     "bwd_GBps_algorithmic": 2 * b * n * 56 / (tb * 1e-3) / 1e9})
    return {"B_global": bg, "n_gpus": world, "scaling": "strong", "rows": rows}


def demo_workload():
    """The reference's only own workload: test_emd() (emd_module.py:81-88): B=20, N=8192, eps=0.002, up to 10000 iterations."""
    b, n, eps, iters = 20, 8192, 0.002, 10000
    g = torch.Generator().manual_seed(0)
    x1, x2 = torch.rand(b, n, 3, generator=g).cuda(), torch.rand(b, n, 3, generator=g).cuda()
    out = {"shape": [b, n, 3], "eps": eps, "iters": iters}
    t = ev_time(lambda: pcl.emd_forward_raw(x1, x2, eps, iters), 3, warm=1)
    d, a, st = pcl.emd_forward_raw(x1, x2, eps, iters, want_stats=True)
    out.update(ours_ms=t, iterations_run=st[:, 1].tolist(), emd=float(d.sqrt().mean()), unique=[int(r.unique().numel()) for r in a])
    try:
        from oracle import build_ref
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
        from helpers import ref_emd_forward
        ref = build_ref.load_ref()
        if ref is not None:
            tr = ev_time(lambda: ref_emd_forward(ref, x1, x2, eps, iters), 1, warm=0)
            rd, ra = ref_emd_forward(ref, x1, x2, eps, iters)
            out.update(reference_ext_ms=tr, reference_emd=float(rd.sqrt().mean()), reference_unique=[int(r.unique().numel()) for r in ra],
                       identical_assignment=[bool(torch.equal(ra[i], a[i])) for i in range(b)])
    except Exception as ex:
        out["reference_ext"] = repr(ex)
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)  # NCCL's banner must not land in the JSON
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier(); torch.cuda.synchronize()
        sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    if args.only == "5" or world > 1:
        res = {"gpu": torch.cuda.get_device_name(local), "config5_chamfer_sweep_B64": config5()}
    else:
        res = {"gpu": torch.cuda.get_device_name(0), "config1_chamfer_cpu_B8_N2048": config1(), "config3_weighted_emd_B32_N2048_C5": config3(),
               "config5_chamfer_sweep_B64": config5(), "reference_demo_B20_N8192": demo_workload()}
    if rank == 0:
        print(json.dumps(res, indent=1))
    if world > 1:
        dist.destroy_process_group()
