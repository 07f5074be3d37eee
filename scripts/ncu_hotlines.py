#!/usr/bin/env python
"""Hot source lines of the kernels in an Nsight Compute report taken with --import-source on (needs -lineinfo).

    python scripts/ncu_hotlines.py report.ncu-rep [kernel-substring] [top_n]

Per kernel: share of executed warp instructions and of stall samples per source line (ncu --page source)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ''
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
cur = fn = hdr = None
per = {}
for r in csv.reader(io.StringIO(raw)):
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) >= 2 and r[0] in ('Function Name', 'Kernel Name'):
        fn = r[1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or len(r) < 10 or r[2] != '-':
        continue
    try:
        line = int(r[0]); inst = int(r[hdr.index('Instructions Executed')]); samp = int(r[hdr.index('# Samples')])
    except ValueError:
        continue
    per.setdefault(fn, []).append((cur, line, r[1][:100], inst, samp))
for fn, agg in per.items():
    if want not in fn:
        continue
    ti, ts = sum(a[3] for a in agg) or 1, sum(a[4] for a in agg) or 1
    print(f'== {fn[:110]}: {ti} warp instructions, {ts} samples')
    for a in sorted(agg, key=lambda a: -a[4])[:top]:
        print(f'{a[0]:16s}{a[1]:5d} inst {100 * a[3] / ti:5.1f}% samples {100 * a[4] / ts:5.1f}%  {a[2]}')
