"""Per-instruction stall samples of an `ncu --page source --csv` dump (first kernel), hot lines only.
usage: ncu_regions.py file.csv [min_exec] [min_samples]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_ex = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
min_s = int(sys.argv[3]) if len(sys.argv) > 3 else 150
hdr = rows[1]
body = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break
    if len(r) == len(hdr):
        body.append(r)
isamp, iex, isrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp]) for r in body)
print('total samples', tot, 'instructions', len(body), 'executed', sum(int(r[iex]) for r in body))
for i in stall:
    t = sum(int(r[i] or 0) for r in body)
    if t > tot * 0.01:
        print('  ', hdr[i], t, f'{100 * t / tot:.1f}%')
for idx, r in enumerate(body):
    ex, s = int(r[iex]), int(r[isamp])
    if ex >= min_ex or s > min_s:
        top = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall), reverse=True)[:1]
        print(f'{idx:4d} {r[isrc].strip()[:58]:58s} ex {ex:>8d} smp {s:>5d} {top[0][1] if s > 100 else ""}')
