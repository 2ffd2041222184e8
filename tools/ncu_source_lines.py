"""Development aid: per-source-line stall samples from `ncu -i REP --page source --print-source cuda,sass --csv`."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; agg = {}
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None or r[0] == "-" or not r[0].isdigit(): continue
    if r[2] != "-": continue  # SASS rows carry an address; keep the per-line aggregate rows only
    k = (cur, int(r[0]))
    s = int(r[hdr["# Samples"]] or 0); ins = int(r[hdr["Instructions Executed"]] or 0)
    a = agg.setdefault(k, [0, 0, r[1].strip()])
    a[0] += s; a[1] += ins
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print(f"total samples {tot}, warp instructions {toti}")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:<5d} {v[0]:7d} {100*v[0]/tot:5.1f}%  instr {100*v[1]/max(toti,1):5.1f}%  {v[2][:120]}")
