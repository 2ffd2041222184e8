#!/usr/bin/env python
"""bench.py -- Chamfer + EMD forward+backward throughput (BASELINE.json metric) on N B200s.

One "step" = one pass of the loss hot path over one batch of B=32 synthetic Table-shaped cloud pairs
(N=M=2048): Chamfer fwd + bwd and auction EMD (eps=0.005, 50 iterations, cfg.py:36-37) fwd + bwd with the
reference's sqrt-mean reduction (utils.py:304, weights == 1).  Weak scaling: every rank owns B=32 clouds.

  value : whole-job clouds/s with inputs resident in HBM, timed with CUDA events, max over ranks
  e2e   : the same metric through the C-ABI host-buffer entry point (pcl_chamfer_emd_step_host): pinned HOST inputs
          copied in, loss scalars copied back and the stream synchronised inside the timed region, every step
          (e2e.python_api: the same through the Python loss classes)
  roofline / cpu_baseline / reference_gpu : see DESIGN.md "Measurement"

`--impl reference` times the reference's CPU path (the oracle port: the reference EMD has no CPU
implementation and pytorch3d is absent) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, NPTS, EPS, ITERS = 32, 2048, 0.005, 50
METRIC = "chamfer+emd fwd+bwd clouds/sec (B=32,N=2048)"
UNIT = "clouds/s"
WORKLOAD = ("config2: Chamfer+EMD fwd+bwd, B=32 per GPU, N=M=2048, Table-shaped synthetic clouds, "
            "regimes independent/noisy alternating, eps=0.005, iters=50")
FLOP_PER_EMD_EVAL = 11   # 8 (distance) + sqrt + 2 adds
FLOP_PER_CHAMFER_EVAL = 8
NCU_AUCTION_DRAM_BYTES = 1670656  # ncu --set full, one launch (profiles/r1_s2_emd_auction_full.txt)


def peaks():
    p = {}
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    return p


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except Exception:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # region shorter than one sample period: take the samples nearest to it
            near = sorted(self.lines, key=lambda tl: min(abs(tl[0] - t0), abs(tl[0] - t1)))[:3]
            for ts, line in near:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except Exception:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ inputs
def make_pool(torch, synth, device, rank, n_sets):
    """n_sets (pred, target) pairs whose total size exceeds the 126 MB L2: 8 seeded base batches (4 'independent'
    = early training, 4 'noisy' = late training) replicated with random point permutations."""
    bases = []
    for s in range(4):
        for regime in ("independent", "noisy"):
            p, t = synth.table_clouds(B_PER_GPU, NPTS, seed=1000 * rank + s, regime=regime)
            bases.append((p.to(device), t[:, :, :3].contiguous().to(device), regime))
    g = torch.Generator(device="cpu").manual_seed(rank)
    pool = []
    for i in range(n_sets):
        p, t, regime = bases[i % len(bases)]
        if i >= len(bases):
            pp = torch.randperm(NPTS, generator=g).to(device)
            tp = torch.randperm(NPTS, generator=g).to(device)
            p, t = p[:, pp].contiguous(), t[:, tp].contiguous()
        pool.append((p, t, regime))
    return pool


class DeviceStep:
    """The hot path through the C ABI with preallocated outputs (no Python allocation in the timed region)."""
    # fill2, chamfer_nn3, chamfer_finish, chamfer_bwd (side stream) | emd_auction, wreduce_stage1, wreduce_stage2, emd_weighted_bwd, emd_mean
    KERNELS_PER_STEP = 9

    def __init__(self, torch, _lib, device):
        self.torch, self.lib, self.L = torch, _lib, _lib.lib()
        b, n = B_PER_GPU, NPTS
        f32, i32 = torch.float32, torch.int32
        e = lambda *s, dt=f32: torch.empty(*s, device=device, dtype=dt)
        self.dist_x, self.dist_y, self.idx_x, self.idx_y = e(b, n), e(b, n), e(b, n, dt=i32), e(b, n, dt=i32)
        self.loss_xy, self.ones = e(2), torch.ones(2, device=device)
        self.gx, self.gy = e(b, n, 3), e(b, n, 3)
        self.dist, self.asg, self.stats = e(b, n), e(b, n, dt=i32), e(b, 8, dt=i32)
        self.sums, self.gemd = e(2), e(b, n, 3)
        self.cws = self.L.pcl_chamfer_workspace_bytes(b, n, n); self.cw = torch.empty(self.cws, device=device, dtype=torch.uint8)
        self.ews = self.L.pcl_emd_workspace_bytes(b, n); self.ew = torch.empty(self.ews, device=device, dtype=torch.uint8)
        self.sws = self.L.pcl_chamfer_emd_step_scratch_bytes(b, n); self.sw = torch.empty(self.sws, device=device, dtype=torch.uint8)
        self.losses = e(4)

    def chamfer(self, p, t, st):
        L, A, b, n = self.L, self.lib.pts_args, B_PER_GPU, NPTS
        rc = L.pcl_chamfer_fwd(*A(p), None, *A(t), None, b, n, n, 3, 0, self.dist_x.data_ptr(), self.idx_x.data_ptr(),
                               self.dist_y.data_ptr(), self.idx_y.data_ptr(), self.loss_xy.data_ptr(), self.cw.data_ptr(), self.cws, st)
        rc |= L.pcl_chamfer_bwd(*A(p), None, *A(t), None, b, n, n, 3, self.idx_x.data_ptr(), self.idx_y.data_ptr(),
                                self.ones.data_ptr(), self.gx.data_ptr(), self.gy.data_ptr(), st)
        return rc

    def emd_fwd(self, p, t, st):
        A, b, n = self.lib.pts_args, B_PER_GPU, NPTS
        return self.L.pcl_emd_fwd(*A(p), *A(t), b, n, EPS, ITERS, self.dist.data_ptr(), self.asg.data_ptr(),
                                  self.stats.data_ptr(), self.ew.data_ptr(), self.ews, st)

    def emd_rest(self, p, t, st):
        L, A, b, n = self.L, self.lib.pts_args, B_PER_GPU, NPTS
        rc = L.pcl_emd_weighted_reduce(self.dist.data_ptr(), None, None, b, n, 0, self.sums.data_ptr(), self.ew.data_ptr(), self.ews, st)
        rc |= L.pcl_emd_weighted_bwd(*A(p), *A(t), b, n, self.asg.data_ptr(), self.dist.data_ptr(), None, None, 0,
                                     self.sums.data_ptr(), self.ones.data_ptr(), self.gemd.data_ptr(), st)
        return rc

    def __call__(self, p, t, st):
        """The whole step in ONE C-ABI call (pcl_chamfer_emd_step): Chamfer on the library's side stream next to the auction."""
        A = self.lib.pts_args
        rc = self.L.pcl_chamfer_emd_step(*A(p), *A(t), B_PER_GPU, NPTS, EPS, ITERS, 0, self.losses.data_ptr(), self.gx.data_ptr(),
                                         self.gemd.data_ptr(), self.sw.data_ptr(), self.sws, st)
        if rc:
            raise RuntimeError(self.L.pcl_last_error().decode())


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_step(oracle, np, x1, x2, threads):
    """The same step on the host cores with the oracle port (Chamfer fwd+bwd, EMD fwd, sqrt-mean, EMD bwd)."""
    c = oracle.chamfer_forward(x1, x2, nthreads=threads)
    oracle.chamfer_backward(x1, x2, c["idx_x"], c["idx_y"], 1.0)
    r = oracle.emd_forward(x1, x2, EPS, ITERS, nthreads=threads)
    d = r["dist"]
    gd = (1.0 / d.size) / (2.0 * np.sqrt(d))
    oracle.emd_backward(x1, x2, r["assignment"], gd.astype(np.float32))
    return r


def cpu_baseline(sample_clouds, steps, warmup):
    import numpy as np
    import oracle
    from pointcloud_b200 import synth
    threads = os.cpu_count() or 1
    half = max(1, sample_clouds // 2)
    pa, ta = synth.table_clouds(half, NPTS, seed=0, regime="independent")
    pb, tb = synth.table_clouds(sample_clouds - half, NPTS, seed=0, regime="noisy") if sample_clouds > half else (pa[:0], ta[:0])
    x1 = np.ascontiguousarray(np.concatenate([pa.numpy(), pb.numpy()]))
    x2 = np.ascontiguousarray(np.concatenate([ta[:, :, :3].numpy(), tb[:, :, :3].numpy()]))
    order = np.argsort(np.concatenate([np.arange(half) * 2, np.arange(sample_clouds - half) * 2 + 1]), kind="stable")
    x1, x2 = np.ascontiguousarray(x1[order]), np.ascontiguousarray(x2[order])  # interleave heavy/light clouds over the threads
    for _ in range(warmup):
        cpu_step(oracle, np, x1, x2, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(oracle, np, x1, x2, threads)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"value": sample_clouds / dt, "unit": UNIT, "cores": min(threads, sample_clouds), "kind": "port",
            "sample": f"{sample_clouds} clouds (half independent, half noisy) of the same workload per step, {steps} steps, "
                      f"oracle C port (oracle/*.c), {threads} host threads available, batch-parallel",
            "ms_per_step": dt * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 16
    # K steps as asked, unless that would take more than about two minutes on this host (one step of the sample is ~0.2 s on
    # 16 cores): then as many as fit; the line reports the number actually timed
    t1 = max(cpu_baseline(sample, 1, 1)["ms_per_step"] * 1e-3, 1e-3)
    steps, warm = max(1, min(args.steps, int(120.0 / t1))), max(1, min(args.warmup, 2))
    cb = cpu_baseline(sample, steps, warm)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference", "config": {"workload": WORKLOAD, "sample_clouds_per_step": sample},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference EMD has no CPU implementation (CUDA only) and pytorch3d is absent: this arm is the CPU oracle port of "
                    "both; the unmodified reference CUDA extension is timed on the GPU in the default arm (key reference_gpu)"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="value leg + breakdown only (short command for ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import pointcloud_b200 as pcl
    from pointcloud_b200 import _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created; stdout must carry only the JSON line,
        # so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    K, W = args.steps, args.warmup
    sampler = ClockSampler(local)
    sampler.start()  # started early: nvidia-smi needs a few hundred ms before its first sample
    n_sets = 96  # 96 * 1.57 MB of inputs = 151 MB > 126 MB L2: every step reads inputs that are not L2-resident
    pool = make_pool(torch, synth, device, rank, n_sets)
    step = DeviceStep(torch, _lib, device)
    st = torch.cuda.current_stream().cuda_stream
    ev = lambda: torch.cuda.Event(enable_timing=True)

    # ---- value: device-resident inputs, whole step, CUDA events on the launching stream ---------------
    for i in range(W):
        step(*pool[i % n_sets][:2], st)
    barrier()
    t_wall0 = time.time()
    e0, e1 = ev(), ev()
    e0.record()
    for i in range(K):
        step(*pool[(W + i) % n_sets][:2], st)
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_per_step = ms_total / K
    value = world * B_PER_GPU * K / (ms_total * 1e-3)

    # ---- per-kernel breakdown on rank 0's stream (explains `value`; same inputs, the phases run one after the other
    #      here, whereas the timed step above overlaps Chamfer with the auction) ----
    phases = {"chamfer_fwd_bwd": [], "emd_fwd": [], "emd_reduce_bwd": []}
    per_regime = {"independent": [], "noisy": []}
    sum_u, executed, seq_ms = [], [], []
    for i in range(min(K, 32)):
        p, t, regime = pool[(W + i) % n_sets]
        a, b_, c, d = ev(), ev(), ev(), ev()
        a.record(); step.chamfer(p, t, st); b_.record(); step.emd_fwd(p, t, st); c.record(); step.emd_rest(p, t, st); d.record()
        torch.cuda.synchronize()
        per_regime_seq = a.elapsed_time(d)
        phases["chamfer_fwd_bwd"].append(a.elapsed_time(b_)); phases["emd_fwd"].append(b_.elapsed_time(c)); phases["emd_reduce_bwd"].append(c.elapsed_time(d))
        e_a, e_b = ev(), ev()
        e_a.record(); step(p, t, st); e_b.record()
        torch.cuda.synchronize()
        per_regime[regime].append(e_a.elapsed_time(e_b))
        seq_ms.append(per_regime_seq)
        if step.chamfer(p, t, st) | step.emd_fwd(p, t, st) | step.emd_rest(p, t, st):
            raise RuntimeError(_lib.lib().pcl_last_error().decode())
        torch.cuda.synchronize()
        sum_u.append(int(step.stats[:, 0].sum().item()))
        executed.append(int(((step.stats[:, 4].long() & 0xffffffff) + (step.stats[:, 5].long() << 32)).sum().item()))
    emd_ms = statistics.mean(phases["emd_fwd"])
    evals = statistics.mean(sum_u) * NPTS  # pair evaluations one auction launch executes (sum_t U_t * N over the batch)
    sm_count = ctypes.c_int(0)
    _lib.lib().pcl_device_info(ctypes.byref(sm_count), None, None, None)
    pk = peaks()
    sm_max = pk.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965.0
    fp32_peak_tflops = sm_count.value * 128 * 2 * sm_max * 1e6 / 1e12
    achieved = FLOP_PER_EMD_EVAL * evals / (emd_ms * 1e-3) / 1e12
    roofline = {"kernel": "emd_auction_kernel", "bound": "fp32-cuda-core", "achieved": achieved, "peak": fp32_peak_tflops,
                "unit": "TFLOP/s", "frac": achieved / fp32_peak_tflops, "traffic": NCU_AUCTION_DRAM_BYTES,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, profiles/r1_s2_emd_auction_full.txt "
                                  "(algorithmic: 1.57 MB inputs + 0.52 MB outputs; the auction state never leaves shared memory)",
                "peak_source": f"{sm_count.value} SMs x 128 lanes x 2 FLOP x {sm_max:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz); "
                               "contraction depth 3 => CUDA-core bound, neither hbm nor tensor (SURVEY.md 8d)",
                "algorithmic": f"{FLOP_PER_EMD_EVAL} FLOP x N x sum_t U_t = {FLOP_PER_EMD_EVAL * evals:.3e} FLOP per launch",
                "pair_evals_per_s": evals / (emd_ms * 1e-3), "avg_launch_ms": emd_ms,
                "executed_fraction": statistics.mean(executed) / evals,
                "executed_note": "algorithmic = every (bidder, target) pair of the reference's Bid (N * sum_t U_t); the kernel proves "
                                 "whole 32-target tiles irrelevant with an exact bounding-box test and really evaluates only this fraction"}
    # secondary roofline: the Chamfer forward kernel alone (FP32 CUDA-core bound as well).  `achieved` counts the ALGORITHMIC 8 FLOP
    # per directed evaluation; the kernel executes 3 FFMA2-halves (6 FLOP) per evaluation in its approximate scan and
    # the exact arithmetic only for the candidate chunks, so `executed_fma_lane_ops_frac` (3 FMA-pipe lane-ops per evaluation against
    # the 128 lanes/clk/SM) is the pipe utilisation the scan itself accounts for; ncu: profiles/r1_s2_chamfer_nn3_full.txt
    tch = []
    for i in range(min(K, 32)):
        p, t, _ = pool[(W + i) % n_sets]
        a, b_ = ev(), ev()
        a.record()
        rc = step.L.pcl_chamfer_fwd(*_lib.pts_args(p), None, *_lib.pts_args(t), None, B_PER_GPU, NPTS, NPTS, 3, 0, step.dist_x.data_ptr(),
                                    step.idx_x.data_ptr(), step.dist_y.data_ptr(), step.idx_y.data_ptr(), step.loss_xy.data_ptr(),
                                    step.cw.data_ptr(), step.cws, st)
        b_.record()
        torch.cuda.synchronize()
        assert rc == 0
        tch.append(a.elapsed_time(b_))
    ch_evals = 2.0 * B_PER_GPU * NPTS * NPTS
    ch_ms = statistics.mean(tch)
    ch_achieved = FLOP_PER_CHAMFER_EVAL * ch_evals / (ch_ms * 1e-3) / 1e12
    roofline_chamfer = {"kernel": "chamfer_nn3_kernel (+memset, finish)", "bound": "fp32-cuda-core", "achieved": ch_achieved, "peak": fp32_peak_tflops,
                        "unit": "TFLOP/s", "frac": ch_achieved / fp32_peak_tflops, "traffic": None, "avg_launch_ms": ch_ms,
                        "algorithmic": f"{FLOP_PER_CHAMFER_EVAL} FLOP x 2*B*N*M directed evaluations = {FLOP_PER_CHAMFER_EVAL * ch_evals:.3e} FLOP per launch",
                        "executed_fma_lane_ops_frac": 3 * ch_evals / (ch_ms * 1e-3) / (fp32_peak_tflops * 1e12 / 2)}
    breakdown = {k: statistics.mean(v) for k, v in phases.items()}
    breakdown["ms_per_step_independent"] = statistics.mean(per_regime["independent"]) if per_regime["independent"] else None
    breakdown["ms_per_step_noisy"] = statistics.mean(per_regime["noisy"]) if per_regime["noisy"] else None
    breakdown["chamfer_directed_pair_evals_per_s"] = ch_evals / (breakdown["chamfer_fwd_bwd"] * 1e-3)
    breakdown["ms_per_step_phases_run_sequentially"] = statistics.mean(seq_ms)

    # ---- e2e: HOST buffers in, HOST results out, every step, through the C ABI (pcl_chamfer_emd_step_host) ------
    # timed region per step: H2D of that step's pinned inputs, all 7 kernels, D2H of the three loss scalars, stream sync
    host = [(p.cpu().pin_memory(), t.cpu().pin_memory()) for p, t, _ in pool[:16]]
    L = _lib.lib()
    nbytes = L.pcl_loss_host_scratch_bytes(B_PER_GPU, NPTS)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=device)
    loss_h = torch.zeros(4).pin_memory()
    cur_stream = torch.cuda.current_stream()

    def host_step(ph, th):
        rc = L.pcl_chamfer_emd_step_host(ph.data_ptr(), th.data_ptr(), B_PER_GPU, NPTS, EPS, ITERS, 0, loss_h.data_ptr(), None, None,
                                         scratch.data_ptr(), nbytes, st)
        if rc:
            raise RuntimeError(L.pcl_last_error().decode())
        cur_stream.synchronize()  # the caller reads loss_h now
        return float(loss_h[0]) + float(loss_h[1]) + float(loss_h[2])

    Ke = max(8, min(K, 200))
    for i in range(3):
        host_step(*host[i % len(host)])
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for i in range(Ke):
        host_step(*host[i % len(host)])
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e = {"value": world * B_PER_GPU * Ke / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 2 * B_PER_GPU * NPTS * 3 * 4,
           "d2h_bytes_per_step": 12, "steps": Ke, "ms_per_step": e2e_ms / Ke,
           "path": "C ABI pcl_chamfer_emd_step_host: pinned host inputs -> H2D -> Chamfer fwd+bwd, EMD fwd, sqrt-mean, EMD bwd -> D2H of the 3 loss "
                   "scalars -> stream sync, every step (gradients stay on the device, as in training)"}

    # ---- the same through the Python loss API (autograd Functions), for the torch user -----------------------------
    emd_mod = pcl.emdModule()

    def api_step(ph, th):
        p = ph.to(device, non_blocking=True).requires_grad_()
        t = th.to(device, non_blocking=True)
        closs, _ = pcl.chamfer_distance(p, t)
        d, _ = emd_mod(p, t, EPS, ITERS)
        eloss = d.sqrt().mean()
        (closs + eloss).backward()
        return torch.stack([closs.detach(), eloss.detach()]).cpu()  # device -> host read of the step's result (synchronises)

    Kp = max(8, min(K, 40))
    for i in range(3):
        api_step(*host[i % len(host)])
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for i in range(Kp):
        api_step(*host[i % len(host)])
    e1.record()
    barrier()
    api_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e["python_api"] = {"value": world * B_PER_GPU * Kp / (api_ms * 1e-3), "ms_per_step": api_ms / Kp, "steps": Kp,
                         "path": "pointcloud_b200.chamfer_distance + emdModule + autograd, pinned host inputs, losses read back"}

    # ---- the unmodified reference CUDA extension on the same GPU (EMD forward only; context, not the target) ----
    reference_gpu = None
    cb = None
    if rank == 0:
        try:
            from oracle import build_ref
            ref = build_ref.load_ref()
            if ref is not None:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                from helpers import ref_emd_forward
                ts = []
                for i in range(6):
                    p, t, _ = pool[i]
                    a, b_ = ev(), ev()
                    a.record(); ref_emd_forward(ref, p, t, EPS, ITERS); b_.record()
                    torch.cuda.synchronize()
                    if i >= 2:
                        ts.append(a.elapsed_time(b_))
                reference_gpu = {"emd_fwd_ms": statistics.mean(ts), "ours_emd_fwd_ms": emd_ms,
                                 "what": "unmodified reference emd extension (oracle/_ref/emd.so, 351 launches) on the same B200, same inputs"}
        except Exception as ex:  # the reference build is optional context
            reference_gpu = {"unavailable": repr(ex)}
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(16, 40, 2)  # ~10 s of CPU work on 16 cores
            cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "clouds_per_gpu": B_PER_GPU, "points": NPTS, "eps": EPS, "iters": ITERS,
                           "chamfer_mode": "unfused", "parallelism": f"batch-sharded x{world}, no data-path collective",
                           "l2": f"inputs rotate through {n_sets} sets = {n_sets * 2 * B_PER_GPU * NPTS * 12 / 1e6:.0f} MB > 126 MB L2 (no flush needed)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": DeviceStep.KERNELS_PER_STEP * K, "roofline": roofline,
                "roofline_chamfer": roofline_chamfer,
                "cpu_baseline": cb, "breakdown_ms": breakdown, "reference_gpu": reference_gpu, "impl": "ours"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
