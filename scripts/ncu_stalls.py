#!/usr/bin/env python
"""Top instructions per stall reason from a .ncu-rep source page."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 8
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name': break
    if len(r) < 10 or r[0] == 'Address': continue
    data.append(r)
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot_all = sum(int(r[ix['# Samples']]) for r in data)
for h in reasons:
    tot = sum(int(r[ix[h]] or 0) for r in data)
    if tot < 0.02 * tot_all: continue
    print(f"== {h}: {tot} samples ({tot/tot_all*100:.1f}% of all)")
    top = sorted(range(len(data)), key=lambda i: -int(data[i][ix[h]] or 0))[:topn]
    for i in top:
        print(f"   {i:5d} {int(data[i][ix[h]])/tot*100:5.1f}%  {data[i][1][:70]}")
