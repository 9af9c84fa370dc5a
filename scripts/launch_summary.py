#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel."""
import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
skip = ("dfma_probe", "FillFunctor", "elementwise")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except Exception: continue
    k = r[ki][:64]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for k, a in agg.items() if not any(s in k for s in skip))
for k, a in agg.items():
    own = not any(s in k for s in skip)
    print(f"{k:66s} n={a[0]:4d} total={a[1]/1e3:10.1f} us  mean={a[1]/a[0]/1e3:8.1f} us" + (f"  share={a[1]/tot*100:5.1f}%" if own else "  (not part of the step)"))
