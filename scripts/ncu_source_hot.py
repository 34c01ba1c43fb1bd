#!/usr/bin/env python
"""Hot SASS regions of one kernel from `ncu -i rep --page source --csv`: top instructions by stall samples."""
import csv
import sys


def main(path, top=45):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    tot = sum(int(r[ix['# Samples']] or 0) for r in body)
    execd = sum(int(r[ix['Instructions Executed']] or 0) for r in body)
    print(f'instructions {len(body)}, samples {tot}, warp-instructions executed {execd}')
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    ranked = sorted(range(len(body)), key=lambda i: -int(body[i][ix['# Samples']] or 0))[:top]
    for i in sorted(ranked):
        r = body[i]
        st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        sts = ' '.join(f'{n}:{v}' for v, n in st if v)
        print(f'{i:5d} {int(r[ix["# Samples"]]):6d} smp {int(r[ix["Instructions Executed"]]):9d} ex  {r[ix["Source"]].strip()[:70]:70s} {sts}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
