#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum per launch, CSV) by kernel name.  usage: summarize_launches.py file.csv [top]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]; ci = {n: i for i, n in enumerate(h)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= ci['Metric Value']:
        continue
    v = float(r[ci['Metric Value']].replace(',', '')); unit = r[ci['Metric Unit']]
    v = v / 1e3 if unit == 'ns' else v * 1e3 if unit == 'ms' else v
    name = re.sub(r'^void ', '', r[ci['Kernel Name']])
    name = re.sub(r'\(.*', '', name)[:70]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f'launches {sum(a[0] for a in agg.values())}, total {tot / 1e3:.2f} ms (cold-cache, serialised: compare shares)')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{t:9.0f} us {100 * t / tot:5.1f}%  n={n:4d}  avg {t / n:8.1f} us  {k}')
