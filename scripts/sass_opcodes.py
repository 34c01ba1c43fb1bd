#!/usr/bin/env python
"""Per-kernel SASS opcode summary of the shipped library (evidence that the hot path is tcgen05 / TMEM / bulk-copy
code): counts of UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / .st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk),
UTMALDG/UTMASTG (tensor-map TMA), SYNCS (mbarrier), FFMA2, LDS / STS, LDG / STG per kernel.

    python scripts/sass_opcodes.py [lib.so] > profiles/r2_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'synergynet_b200', 'libsynergy_b200.so')
OPS = ['UTCHMMA', 'LDTM', 'STTM', 'UTCBAR', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'FFMA2', 'FFMA', 'LDS', 'STS', 'LDG', 'STG',
       'ATOM', 'RED', 'BAR']
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
counts, cur, total = collections.OrderedDict(), None, 0
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        op = m.group(1)
        counts[cur]['_total'] += 1
        for o in OPS:
            if op == o or op.startswith(o + '.') or (o in ('LDS', 'STS', 'LDG', 'STG', 'ATOM', 'RED', 'BAR', 'SYNCS') and op.startswith(o)):
                counts[cur][o] += 1
                break
print(f'# {os.path.basename(lib)}: SASS instruction counts per kernel (cuobjdump -sass, sm_100a)')
print('# ' + ' '.join(f'{o:>7s}' for o in ['total'] + OPS) + '  kernel')
tot = collections.Counter()
for k, c in counts.items():
    name = demangle(k)
    name = re.sub(r'syn::', '', name)
    name = re.sub(r'\(.*', '', name)
    print('  ' + ' '.join(f'{c[o]:7d}' for o in ['_total'] + OPS) + '  ' + name[:150])
    tot.update(c)
print('  ' + ' '.join(f'{tot[o]:7d}' for o in ['_total'] + OPS) + '  ALL KERNELS')
