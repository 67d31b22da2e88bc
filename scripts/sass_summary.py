"""Counts of the Blackwell-specific SASS instructions per compiled object of libdhd_b200.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA tensor loads / stores), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk),
UTCBAR (tcgen05.commit), SYNCS (mbarrier).  Usage: python scripts/sass_summary.py > profiles/sass_summary.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r'\b(UTCHMMA(?:\.2CTA)?|UTMALDG(?:\.\dD)?(?:\.2CTA)?|UTMASTG(?:\.\dD)?|LDTM(?:\.x\d+)?|UBLKCP(?:\.[A-Z.]+)?|UTCBAR(?:\.2CTA)?(?:\.MULTICAST)?|UTCATOMSWS(?:\.2CTA)?|SYNCS\.[A-Z]+|MUFU\.[A-Z0-9]+)')


def main():
    print('# SASS evidence (sm_100a) per object of dhd_b200/libdhd_b200.so -- made by scripts/sass_summary.py')
    for obj in sorted(glob.glob(os.path.join(ROOT, 'dhd_b200', 'build', '*.cu.o'))):
        sass = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
        per_fn, fn = collections.OrderedDict(), None
        for line in sass.splitlines():
            m = re.match(r'\s*Function : (\S+)', line)
            if m:
                fn = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
                per_fn[fn] = collections.Counter()
                continue
            if fn is not None:
                for tok in PAT.findall(line):
                    per_fn[fn][tok] += 1
        print('\n== %s' % os.path.basename(obj))
        for fn, c in per_fn.items():
            keys = [k for k in c if not k.startswith(('SYNCS', 'MUFU'))]
            if not keys:
                continue
            print('  %-60s %s' % (fn[:60], ', '.join('%s x%d' % (k, c[k]) for k in sorted(keys))))


if __name__ == '__main__':
    main()
