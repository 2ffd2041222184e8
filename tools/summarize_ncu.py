"""Turn gpurun_out/*.ncu-rep / launch-list CSVs into the small text summaries committed under profiles/."""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__cluster_dim_x", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
        res.append(d)
    return res


def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = defaultdict(int)
    samples = 0
    for r in rows[2:]:
        if len(r) < len(hdr) or r[ci["# Samples"]] == "# Samples":  # short rows / repeated headers of further launches
            continue
        samples += int(r[ci["# Samples"]] or 0)
        for s in names:
            tot[s] += int(r[ci[s]] or 0)
    return samples, sorted(tot.items(), key=lambda kv: -kv[1])


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        name = r[ki].split("(")[0]
        tot[name] += v
        cnt[name] += 1
    s = sum(tot.values())
    lines = [f"{'ms':>10} {'share':>6} {'n':>5}  kernel"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        lines.append(f"{v:10.3f} {100 * v / s:5.1f}% {cnt[k]:5d}  {k[:110]}")
    return "\n".join(lines)


if __name__ == "__main__":
    kind, path = sys.argv[1], sys.argv[2]
    if kind == "launches":
        print(launches(path))
    else:
        for d in raw(path):
            print("kernel:", d.get("Kernel Name", ("?",))[0][:120])
            for k in KEYS:
                if k in d:
                    print(f"  {k:72s} {d[k][0]:>16s} {d[k][1]}")
        n, st = stalls(path)
        print(f"warp stall samples: {n}")
        for k, v in st[:10]:
            print(f"  {k:28s} {v:8d} {100 * v / max(n, 1):5.1f}%")
