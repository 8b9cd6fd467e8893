#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) > mv:
        agg[r[kn][:70]].append(float(r[mv].replace(',', '')))
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:70s} n={len(v):4d} total={sum(v)/1e3:10.1f}us avg={sum(v)/len(v)/1e3:9.2f}us share={100*sum(v)/tot:5.1f}%")
