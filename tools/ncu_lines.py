"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line (dev tool).
usage: ncu_lines.py export.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
hdr = None
agg = {}
inlined = {}
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[2] != "-":
        continue  # SASS row
    line = int(r[0])
    d = dict(zip(hdr[4:], r[4:]))
    a = agg.setdefault(line, dict(src=r[1], samples=0, inst=0, tinst=0, shw=0, shideal=0))
    a["samples"] += int(d["# Samples"]); a["inst"] += int(d["Instructions Executed"]); a["tinst"] += int(d["Thread Instructions Executed"])
    a["shw"] += int(d["L1 Wavefronts Shared"]); a["shideal"] += int(d["L1 Wavefronts Shared Ideal"])
tot_i = sum(a["inst"] for a in agg.values()); tot_s = sum(a["samples"] for a in agg.values()); tot_w = sum(a["shw"] for a in agg.values())
print(f"total warp-inst {tot_i}  samples {tot_s}  shared wavefronts {tot_w}")
print(f"{'line':>5} {'inst%':>6} {'smp%':>6} {'lanes':>5} {'shw%':>5}  source")
for line, a in sorted(agg.items(), key=lambda kv: -kv[1]["inst"])[:top]:
    if a["inst"] == 0: continue
    print(f"{line:5d} {100*a['inst']/tot_i:6.2f} {100*a['samples']/max(tot_s,1):6.2f} {a['tinst']/a['inst']:5.1f} {100*a['shw']/max(tot_w,1):5.1f}  {a['src'].strip()[:110]}")
