"""Print an ncu launch list (gpu__time_duration.sum CSV) as one line per launch, plus totals per kernel.
Usage: python scripts/launch_list.py file.csv [--agg]"""
import csv
import re
import sys
import collections

f = sys.argv[1]
lines = [l for l in open(f) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, gi, si, ui = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Grid Size', 'Stream', 'Metric Unit'))
tot = 0.0
agg = collections.defaultdict(lambda: [0, 0.0])
for i, row in enumerate(r):
    v = float(row[vi].replace(',', ''))
    v = v / 1000 if row[ui] == 'ns' else (v * 1000 if row[ui] == 'ms' else v)
    tot += v
    name = re.sub(r'\(.*', '', row[ki])
    name = re.sub(r'^void ', '', name)[:56]
    agg[name][0] += 1
    agg[name][1] += v
    if '--agg' not in sys.argv:
        print('%4d s%-3s %-56s %-16s %8.1f' % (i, row[si], name, row[gi], v))
print('total %.1f us' % tot)
if '--agg' in sys.argv:
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print('%9.1f us %5.1f%%  n=%4d  %s' % (v[1], 100 * v[1] / tot, v[0], k))
