"""Summarise an ncu report (made with --set full --import-source on) into a small text file:
key raw metrics per captured launch + stall-reason totals + the hottest SASS lines.
Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__cycles_active.avg', 'smsp__inst_executed.sum']


def run(args):
    return subprocess.run(['ncu', '-i'] + args, capture_output=True, text=True).stdout


def main(rep, out):
    lines = []
    rows = list(csv.reader(io.StringIO(run([rep, '--page', 'raw', '--csv']))))
    h = rows[0]
    for r in rows[2:]:
        name = r[h.index('Kernel Name')] if 'Kernel Name' in h else '?'
        lines.append('== launch: %s' % name[:100])
        for k in KEYS:
            if k in h:
                lines.append('  %-64s %12s %s' % (k, r[h.index(k)], rows[1][h.index(k)]))
        if 'dram__bytes_read.sum' in h:
            def val(k):
                v, u = float(r[h.index(k)]), rows[1][h.index(k)]
                return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
            lines.append('  dram traffic per launch (read+write)                            %12.0f byte'
                         % (val('dram__bytes_read.sum') + val('dram__bytes_write.sum')))
    src = list(csv.reader(io.StringIO(run([rep, '--page', 'source', '--csv']))))
    hi = [i for i, r in enumerate(src) if r and r[0] == 'Address']
    if hi:
        h = src[hi[0]]
        end = hi[1] - 1 if len(hi) > 1 else len(src)
        body = [r for r in src[hi[0] + 1:end] if len(r) == len(h)]
        ix = {k: i for i, k in enumerate(h)}
        tot = sum(int(r[ix['# Samples']] or 0) for r in body) or 1
        stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
        agg = {k: sum(int(r[ix[k]] or 0) for r in body) for k in stalls}
        lines.append('== warp-stall samples of the first launch (total %d)' % tot)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
            lines.append('  %-28s %7d  %5.1f%%' % (k, v, 100.0 * v / tot))
        lines.append('  warp instructions executed   %d' %
                     sum(int(r[ix['Instructions Executed']] or 0) for r in body))
        lines.append('== hottest SASS lines (samples, executed, instruction)')
        for r in sorted(body, key=lambda r: -int(r[ix['# Samples']] or 0))[:15]:
            lines.append('  %6s %9s  %s' % (r[ix['# Samples']], r[ix['Instructions Executed']],
                                            r[ix['Source']][:100]))
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
