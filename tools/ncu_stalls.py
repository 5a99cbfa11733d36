"""Stall-reason totals (and per source line, top N) from an
`ncu -i rep --page source --csv --print-source cuda,sass` export.
usage: python tools/ncu_stalls.py export.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
cur = None
tot = collections.Counter()
per_line = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        hdr = None
        continue
    if r and r[0] == 'Line No':
        hdr = r
        cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        isamp = hdr.index('# Samples')
        continue
    if hdr is None or not r or r[0] == '':
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    if cur is None:
        continue
    c = collections.Counter()
    for i, h in cols:
        try:
            v = int(r[i] or 0)
        except (ValueError, IndexError):
            v = 0
        if v:
            c[h] += v
    if c:
        tot.update(c)
        per_line[(cur, ln)] = (c, r[1])
n = sum(tot.values()) or 1
print('stall samples by reason (%d total):' % n)
for h, v in tot.most_common():
    print('  %-26s %5.1f%%' % (h, 100.0 * v / n))
print('top lines:')
for k, (c, src) in sorted(per_line.items(), key=lambda kv: -sum(kv[1][0].values()))[:top]:
    s = sum(c.values())
    why = ' '.join('%s=%.1f' % (h.replace('stall_', ''), 100.0 * v / n) for h, v in c.most_common(3))
    print('%-18s %4d %5.1f%%  [%s]  %s' % (k[0], k[1], 100.0 * s / n, why, src.strip()[:70]))
