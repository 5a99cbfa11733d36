"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line.
usage: python tools/ncu_lines.py file.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cur = None
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        hdr = None
        continue
    if r and r[0] == 'Line No':
        hdr = r
        ie = hdr.index('Instructions Executed')
        isamp = hdr.index('# Samples')
        continue
    if hdr is None or len(r) <= ie or r[0] == '':
        continue
    try:
        ln = int(r[0])
        inst = int(r[ie] or 0)
        samp = int(r[isamp] or 0)
    except ValueError:
        continue
    agg[(cur, ln)] = (inst, samp, r[1])
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print('total warp instructions %d, stall samples %d' % (tot, tots))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-20s %4d  inst %5.1f%%  samples %5.1f%%  %s" % (k[0], k[1], 100 * v[0] / tot, 100 * v[1] / tots, v[2][:88]))
