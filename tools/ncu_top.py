"""Top stall-sampled SASS instructions of an ncu report (source page):  python tools/ncu_top.py report.ncu-rep [n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hdr]
si, ci = h.index('Source'), h.index('# Samples')
body = [r for r in rows[hdr + 1:] if len(r) > ci]
tot = sum(int(r[ci] or 0) for r in body)
print('total samples', tot)
for idx, r in sorted(enumerate(body), key=lambda t: -int(t[1][ci] or 0))[:n]:
    print('%5d %6.2f%%  #%4d  %s' % (int(r[ci]), 100.0 * int(r[ci]) / max(tot, 1), idx, r[si].strip()))
