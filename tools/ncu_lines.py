"""Summarise an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line:
share of executed warp instructions and of stall samples.  usage: ncu_lines.py dump.csv [nblocks]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = None
lines = collections.OrderedDict()
fname = ""
def f(x):
    try: return float(x)
    except ValueError: return 0.0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r; si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) != len(hdr) or r[0] == "": continue
    try: ln = int(r[0])
    except ValueError: continue
    key = (fname, ln, r[1].strip()[:100])
    v = lines.setdefault(key, [0.0, 0.0])
    v[0] += f(r[si]); v[1] += f(r[ie])
tot_s = sum(v[0] for v in lines.values()); tot_i = sum(v[1] for v in lines.values())
print("samples %d  warp-instructions %d  per unit %.0f" % (tot_s, tot_i, tot_i / units))
for (fn, ln, src), (s, i) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:50]:
    print("%s:%d %-82s inst %5.1f%% samp %5.1f%%" % (fn[:14], ln, src[:82], 100 * i / tot_i, 100 * s / tot_s))
