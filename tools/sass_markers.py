#!/usr/bin/env python
"""Blackwell instruction markers per compiled object (cuobjdump -sass | count): tcgen05 MMA (UTCHMMA / UTCQMMA ...), tcgen05.cp
(UTCCP), TMEM loads / stores (LDTM / STTM), TMA tensor loads / stores (UTMALDG / UTMASTG), bulk copies (UBLKCP), bulk
reductions (UTMAREDG), mbarrier waits (SYNCS), 256-bit global accesses (LDG / STG .256).  Run after build(); no GPU needed."""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARKERS = ['UTCHMMA', 'UTCQMMA', 'UTCOMMA', 'UTCCP', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UBLKCP', 'UTMAREDG',
           'SYNCS', 'LDG.E.256', 'STG.E.256', 'REDG.E', 'HMMA', 'FFMA']


def main():
    print('# cuobjdump -sass marker counts per object (sm_100a); kernels = functions in the object')
    print('%-22s %s' % ('object', ' '.join('%9s' % m for m in MARKERS)))
    for obj in sorted(glob.glob(os.path.join(ROOT, 'build', '*.o'))):
        sass = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
        counts = collections.Counter()
        for line in sass.splitlines():
            for m in MARKERS:
                if re.search(r'\b' + re.escape(m), line):
                    counts[m] += 1
        print('%-22s %s' % (os.path.basename(obj), ' '.join('%9d' % counts[m] for m in MARKERS)))
    sys.stdout.flush()


if __name__ == '__main__':
    main()
