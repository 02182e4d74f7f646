#!/usr/bin/env python
"""Per-kernel totals from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, sys
path, nsteps = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 2:]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
skip = sys.argv[3].split(',') if len(sys.argv) > 3 else ['k_sim_']   # set-up kernels (data generation) are not part of a step
for r in data:
    if len(r) <= vi or any(x in r[ki] for x in skip): continue
    a = agg.setdefault(r[ki][:64], [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(',', ''))
tot = sum(a[1] for a in agg.values())
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n:64s} n={a[0]/nsteps:6.1f}/step {a[1]/nsteps/1e3:9.1f} us/step {a[1]/a[0]/1e3:8.1f} us each {100*a[1]/tot:5.1f}%")
print('total us/step', tot / nsteps / 1e3)
