#!/usr/bin/env python
"""Top source lines of an `ncu --page source --csv --print-source cuda,sass` export, per kernel: executed warp instructions and stall samples.
usage: tools/ncu_source_top.py source.csv[.gz] [kernel-substring] [N]"""
import csv, gzip, sys, collections, io
path = sys.argv[1]; filt = sys.argv[2] if len(sys.argv) > 2 else ""; N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = csv.reader(f)
kern = None; fpath = None; hdr = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": kern = r[1].split("(")[0].replace("void b200c::", ""); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or kern is None: continue
    if r[0] == "": continue  # SASS rows (attributed to the preceding source line already)
    d = dict(zip(hdr, r))
    try:
        inst = int(d.get("Instructions Executed", 0) or 0); samp = int(d.get("# Samples", 0) or 0)
    except ValueError:
        continue
    k = agg.setdefault(kern, collections.OrderedDict())
    key = (fpath, int(r[0]))
    e = k.setdefault(key, [0, 0, r[1].strip()[:110], 0, 0])
    e[0] += inst; e[1] += samp
    try:
        e[3] += int(d.get("L1 Wavefronts Shared Excessive", 0) or 0); e[4] += int(d.get("L2 Theoretical Sectors Global Excessive", 0) or 0)
    except ValueError:
        pass
for kern, k in agg.items():
    if filt not in kern: continue
    ti = sum(v[0] for v in k.values()); ts = sum(v[1] for v in k.values())
    print(f"==== {kern}: {ti/1e6:.1f} M warp-instr, {ts} samples")
    for (fp, ln), v in sorted(k.items(), key=lambda kv: -kv[1][0])[:N]:
        print(f"{100*v[0]/max(ti,1):5.1f}% inst {100*v[1]/max(ts,1):5.1f}% smp  shx={v[3]:>9} {fp}:{ln}: {v[2]}")
