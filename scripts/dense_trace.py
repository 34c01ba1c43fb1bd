#!/usr/bin/env python
"""Timeline of dense_recon_fm_kernel's CTA 0 (SYN_DENSE_TRACE=<file> makes the library dump clock64 stamps):
per item, epilogue group: [wait alpha+meta | wait MMAs | TMEM round 0 | round 1 incl. stores of round 0 | stores] and
issuer: [wait alpha | wait TMEM buffer | plane 0 / 1 / 2 landed | MMAs issued | alpha prefetch]."""
import sys

rows = [l.split() for l in open(sys.argv[1])]
ep = {(r[0], int(r[1])): [int(x) for x in r[2:]] for r in rows}
t0 = min(v[0] for v in ep.values() if v[0] > 0)
print(' item | issuer: start  w.alpha  w.tmem  plane0  plane1  plane2  issued  prefetch | epilogue: start  w.meta  w.mma   ld0    ld1  stores   end')
for i in range(0, 46):
    iss = ep.get(('issuer', i)); e = ep.get(('epi%d' % (i & 1), i))
    if not iss or iss[0] == 0:
        break
    d = lambda a, k: a[k] - a[k - 1] if a[k] and a[k - 1] else -1
    print(f' {i:4d} | {iss[0] - t0:13d} {d(iss, 1):8d} {d(iss, 2):7d} {d(iss, 3):7d} {d(iss, 4):7d} {d(iss, 5):7d} {d(iss, 6):7d} {d(iss, 7):9d} |'
          f' {e[0] - t0:15d} {d(e, 1):7d} {d(e, 2):6d} {d(e, 3):6d} {d(e, 4):6d} {d(e, 5):7d} {e[5] - t0:6d}')
