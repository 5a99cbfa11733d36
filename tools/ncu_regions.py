"""Instruction share per function of search_kernel.cuh from an ncu cuda,sass source export.
usage: python tools/ncu_regions.py export.csv n_queries n_exp_per_query"""
import collections
import csv
import re
import sys

src = open('parallel_hnsw_b200/csrc/search_kernel.cuh').read().split('\n')
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r'\s*__device__ .*?(\w+)\(', l) or re.match(r'\s*__global__ .*', l)
    if m and ('{' in l or l.rstrip().endswith(',') or l.rstrip().endswith('(')):
        name = m.group(1) if m.lastindex else 'kernel'
        starts.append((i, name))
starts.append((len(src) + 1, 'end'))
rows = list(csv.reader(open(sys.argv[1])))
nq = float(sys.argv[2])
nexp = float(sys.argv[3])
cur = None
hdr = None
agg = collections.OrderedDict()
other = 0
tot = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        hdr = None
        continue
    if r and r[0] == 'Line No':
        hdr = r
        ie = hdr.index('Instructions Executed')
        continue
    if hdr is None or len(r) <= ie or r[0] == '':
        continue
    try:
        ln = int(r[0])
        inst = int(r[ie] or 0)
    except ValueError:
        continue
    tot += inst
    if cur != 'search_kernel.cuh':
        other += inst
        continue
    name = 'preamble'
    for (a, n), (b, _) in zip(starts, starts[1:]):
        if a <= ln < b:
            name = n
            break
    agg[name] = agg.get(name, 0) + inst
print("total %.0f instr/query, %.0f per expansion" % (tot / nq, tot / nq / nexp))
for n, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print("%-24s %5.1f%%  %6.0f /expansion" % (n, 100 * v / tot, v / nq / nexp))
print("%-24s %5.1f%%  %6.0f /expansion" % ('(inlined headers)', 100 * other / tot, other / nq / nexp))
