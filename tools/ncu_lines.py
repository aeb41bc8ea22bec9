"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line (dev tool).
usage: ncu_lines.py export.csv [top_n]"""
import csv, sys, os
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
hdr = None
fname = "?"
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    line = int(r[0])
    d = dict(zip(hdr[4:], r[4:]))
    a = agg.setdefault((fname, line), dict(src=r[1], samples=0, inst=0, tinst=0, shw=0, stalls={}))
    a["samples"] += int(d["# Samples"]); a["inst"] += int(d["Instructions Executed"]); a["tinst"] += int(d["Thread Instructions Executed"])
    a["shw"] += int(d["L1 Wavefronts Shared"])
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            a["stalls"][k[6:]] = a["stalls"].get(k[6:], 0) + int(v)
tot_i = sum(a["inst"] for a in agg.values()); tot_s = sum(a["samples"] for a in agg.values()); tot_w = sum(a["shw"] for a in agg.values())
tot_t = sum(a["tinst"] for a in agg.values())
print(f"total warp-inst {tot_i}  thread-inst/warp-inst {tot_t/max(tot_i,1):.1f}  samples {tot_s}  shared wavefronts {tot_w}")
st = {}
for a in agg.values():
    for k, v in a["stalls"].items(): st[k] = st.get(k, 0) + v
print("stalls: " + " ".join(f"{k}={100*v/max(tot_s,1):.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
print(f"{'file:line':>22} {'inst%':>6} {'smp%':>6} {'lanes':>5} {'shw%':>5} top-stall  source")
for (f, line), a in sorted(agg.items(), key=lambda kv: -kv[1]["inst"])[:top]:
    if a["inst"] == 0: continue
    ts = max(a["stalls"].items(), key=lambda kv: kv[1])[0] if a["stalls"] else ""
    print(f"{f[:16]+':'+str(line):>22} {100*a['inst']/tot_i:6.2f} {100*a['samples']/max(tot_s,1):6.2f} {a['tinst']/a['inst']:5.1f} {100*a['shw']/max(tot_w,1):5.1f} {ts:>9}  {a['src'].strip()[:90]}")
