"""Summarise an `ncu --page source --csv` dump: hottest SASS instructions by warp-stall samples (first kernel)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.008
body = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break
    body.append(r)
tot = sum(int(r[2]) for r in body if r[2].isdigit())
print(rows[0][1][:80], 'total samples', tot, 'instructions', len(body))
for i, r in enumerate(body):
    s = int(r[2]) if r[2].isdigit() else 0
    if s > tot * thr:
        print(f'{i:4d} {r[1].strip()[:72]:72s} {s:6d} {100 * s / tot:5.1f}%  exec {r[5]}')
