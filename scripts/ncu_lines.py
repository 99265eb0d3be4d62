"""Per-source-line instruction / stall-sample shares from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name K`.
usage: ncu_lines.py export.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
hdr = None
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_smp = hdr.index("# Samples")
        i_wf = hdr.index("L1 Wavefronts Shared")
        continue
    if hdr is None or len(r) <= i_inst or r[2] != "-":   # source-level rows have "-" as address
        continue
    try:
        ie = float(r[i_inst]); sm = float(r[i_smp]); wf = float(r[i_wf] or 0)
    except ValueError:
        continue
    k = (cur, r[0])
    a = agg.setdefault(k, [0.0, 0.0, 0.0, r[1]])
    a[0] += ie; a[1] += sm; a[2] += wf
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values()); tw = sum(a[2] for a in agg.values())
print(f"total warp inst {ti:.4g}, samples {ts:.4g}, smem wavefronts {tw:.4g}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][int(sys.argv[3]) if len(sys.argv) > 3 else 1])[:top]:
    print(f"{a[1] / ts * 100:5.1f}% smp {a[0] / ti * 100:5.1f}% inst {a[2] / max(tw, 1) * 100:5.1f}% wf  {f}:{ln:>4}  {a[3][:110]}")
