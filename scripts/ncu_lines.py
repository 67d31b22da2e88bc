"""Aggregate an ncu report's source page by CUDA source line (needs -lineinfo + --import-source on).
Usage: python scripts/ncu_lines.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, agg = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = r
        ix = {k: i for i, k in enumerate(hdr)}
    elif r[0] == 'Function Name':
        print('##', r[1][:100])
    elif hdr and r[0].isdigit():
        g = lambda k: int(r[ix[k]] or 0) if k in ix and r[ix[k]].lstrip('-').isdigit() else 0
        agg.append((g('# Samples'), g('Instructions Executed'), g('stall_long_sb'), g('stall_short_sb'),
                    g('stall_lg'), g('stall_mio'), fname, r[0], r[1].strip()[:90]))
ts = sum(a[0] for a in agg) or 1
ti = sum(a[1] for a in agg) or 1
print('total samples %d, warp instructions %d' % (ts, ti))
print('%7s %6s %10s %6s | long short lg mio | line' % ('samples', '%', 'inst', '%'))
for a in sorted(agg, key=lambda a: -a[0])[:top]:
    print('%7d %5.1f%% %10d %5.1f%% | %5d %5d %4d %4d | %s:%s  %s' %
          (a[0], 100.0 * a[0] / ts, a[1], 100.0 * a[1] / ti, a[2], a[3], a[4], a[5], a[6], a[7], a[8]))
