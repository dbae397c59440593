#!/usr/bin/env python
"""Where the warp-stall samples of a kernel fall: python scripts/ncu_source_shares.py x_source.csv.gz [kernel#]"""
import collections, csv, gzip, io, sys
lines = gzip.open(sys.argv[1], 'rt').read().split('\n')
kern = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blocks, cur = [], None
for ln in lines:
    if ln.startswith('"Kernel Name"'):
        cur = {'name': ln[:120], 'hdr': None, 'rows': []}
        blocks.append(cur)
        continue
    if cur is None or not ln:
        continue
    r = next(csv.reader(io.StringIO(ln)))
    if r[0] == 'Address':
        cur['hdr'] = r
    elif r[0].startswith('0x'):
        cur['rows'].append(r)
b = blocks[kern]
print(b['name'])
data = b['rows']
tot = sum(int(r[4]) for r in data)
n = len(data)
print('instructions', n, 'samples', tot)
cls, ex = collections.Counter(), collections.Counter()
for r in data:
    t = r[1].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    cls[op] += int(r[4]); ex[op] += int(r[5])
print('by opcode (share of samples, executed):', ', '.join(f'{k} {100*v/tot:.1f}% ({ex[k]})' for k, v in cls.most_common(14)))
seg = 40
for d in range(seg):
    lo, hi = d * n // seg, (d + 1) * n // seg
    sh = sum(int(r[4]) for r in data[lo:hi]) / tot
    exs = sum(int(r[5]) for r in data[lo:hi])
    if sh > 0.01:
        ops = collections.Counter()
        for r in data[lo:hi]:
            t = r[1].split(); ops[(t[1] if t[0].startswith('@') else t[0]).split('.')[0]] += int(r[5])
        print(f'  instr {lo:6d}-{hi:6d}: {100*sh:5.1f}% samples, {exs:>12} exec, top ops {ops.most_common(4)}')
