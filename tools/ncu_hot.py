#!/usr/bin/env python
"""Hottest SASS lines (by stall samples) of the k-th kernel of an .ncu-rep: ncu_hot.py rep k [nlines]"""
import csv, io, subprocess, sys

def I(x):
    try: return int(float(str(x).replace(',', '') or 0))
    except ValueError: return 0

rep, k = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(k), '--launch-count', '1'],
                     capture_output=True, text=True).stdout
src = list(csv.reader(io.StringIO(out)))
print(src[0][1][:120])
h2 = src[1]
isrc, ins, isamp = h2.index('Source'), h2.index('Instructions Executed'), h2.index('# Samples')
data = [r for r in src[2:] if len(r) > max(isrc, ins, isamp)]
tsamp = sum(I(r[isamp]) for r in data)
stall_cols = [i for i, nm in enumerate(h2) if nm.startswith('stall_') and 'Not Issued' not in nm]
print('total samples', tsamp, 'warp instr', sum(I(r[ins]) for r in data))
order = sorted(range(len(data)), key=lambda i: -I(data[i][isamp]))[:n]
for i in sorted(order):
    r = data[i]
    st = sorted(((I(r[j]), h2[j]) for j in stall_cols if r[j] not in ('', '0')), reverse=True)[:2]
    print('%4d %5.1f%% exec %8s  %-84s %s' % (i, 100 * I(r[isamp]) / max(tsamp, 1), r[ins], r[isrc].strip()[:84], st))
