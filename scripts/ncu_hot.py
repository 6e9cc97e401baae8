#!/usr/bin/env python
"""Per-instruction hot spots from a .ncu-rep source page: prints SASS lines with their share of executed
warp-instructions and of stall samples (only lines above a threshold, plus region markers)."""
import csv, subprocess, sys, io
rep, thresh = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name': break
    if len(r) < 10 or r[0] == 'Address': continue
    data.append(r)
tot = sum(int(r[ix['Instructions Executed']]) for r in data)
samp = sum(int(r[ix['# Samples']]) for r in data)
print('instructions', len(data), 'executed', tot, 'samples', samp)
for i, r in enumerate(data):
    if i < lo or i > hi: continue
    ie = int(r[ix['Instructions Executed']]); te = int(r[ix['Thread Instructions Executed']]); s = int(r[ix['# Samples']])
    if ie / tot * 100 >= thresh or s / samp * 100 >= thresh:
        print(f"{i:5d} {ie/tot*100:5.2f}% thr {te/max(ie,1):4.1f} smp {s/samp*100:5.2f}%  {r[1][:80]}")
