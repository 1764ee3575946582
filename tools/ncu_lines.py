#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.
usage: ncu_lines.py export.csv [topN]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path, errors="replace")))
agg = collections.OrderedDict()
cur_file = None; hdr = None; first_kernel_done = False; kernels = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr, r))   # line-total rows carry the line number; SASS rows have an empty first column
    key = (cur_file, ln)
    def num(v):
        try: return float(v)
        except (TypeError, ValueError): return 0.0
    samp = num(d.get("# Samples")); ins = num(d.get("Instructions Executed"))
    a = agg.setdefault(key, [0.0, 0.0, r[1][:110], collections.Counter()])
    a[0] += samp; a[1] += ins
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v:
            try: a[3][k] += float(v)
            except ValueError: pass
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print("total samples %.0f, total warp-instructions %.0f" % (tot_s, tot_i))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ",".join("%s=%.0f" % (k[6:], v) for k, v in a[3].most_common(3))
    print("%5.1f%% samp %5.1f%% inst  %s:%d  %s   [%s]" % (100 * a[0] / max(tot_s, 1), 100 * a[1] / max(tot_i, 1), f, ln, a[2].strip(), st))
