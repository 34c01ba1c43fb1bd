#!/usr/bin/env python
"""Shared-memory wavefronts per SASS instruction from `ncu -i rep --page source --csv` (SASS view):
which loads/stores carry the wavefronts and which of them are excessive (bank conflicts)."""
import csv
import subprocess
import sys


def main(rep, top=40):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    num = lambda r, k: int(r[ix[k]] or 0)
    tot = sum(num(r, 'L1 Wavefronts Shared') for r in body)
    exc = sum(num(r, 'L1 Wavefronts Shared Excessive') for r in body)
    print(f'shared wavefronts {tot}, excessive {exc} ({100.0 * exc / max(tot, 1):.1f} %)')
    ranked = sorted(range(len(body)), key=lambda i: -num(body[i], 'L1 Wavefronts Shared'))[:top]
    for i in sorted(ranked):
        r = body[i]
        print(f'{i:5d} wf {num(r, "L1 Wavefronts Shared"):9d} excess {num(r, "L1 Wavefronts Shared Excessive"):9d} '
              f'ex {num(r, "Instructions Executed"):8d} smp {num(r, "# Samples"):5d}  {r[ix["Source"]].strip()[:80]}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
