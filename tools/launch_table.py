"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/launch_table.py file.csv [n_steps]"""
import collections
import csv
import io
import re
import sys

txt = open(sys.argv[1]).read()
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.DictReader(io.StringIO('\n'.join(l for l in txt.split('\n') if l.startswith('"')))))
agg, tot = collections.OrderedDict(), 0.0
for r in rows:
    name = re.sub(r'\(.*', '', r['Kernel Name'])[:70]
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print(f'{"us/step":>10}  {"share":>6}  {"launches/step":>13}  kernel')
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{v / steps:10.1f}  {v / tot:6.1%}  {c / steps:13.1f}  {k}')
print(f'{tot / steps:10.1f}  total us/step over {len(rows)} launches')
