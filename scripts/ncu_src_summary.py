#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: stall samples by SASS opcode and by stall reason."""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
print(len(data), 'instrs; total samples', tot)
op = collections.Counter(); cnt = collections.Counter(); ex = collections.Counter()
for r in data:
    s = r[ix['Source']].strip()
    s = re.sub(r'^@!?U?P\d+\s+', '', s)
    o = s.split()[0]
    o = '.'.join(o.split('.')[:2]) if o.startswith(('LD', 'ST')) else o.split('.')[0]
    op[o] += int(r[ix['# Samples']]); cnt[o] += 1; ex[o] += int(r[ix['Instructions Executed']])
te = sum(ex.values())
print('warp instr executed', te)
for o, c in op.most_common(30):
    print(f'{o:12s} samples {c:6d} {100*c/tot:5.1f}%  static {cnt[o]:5d}  exec {100*ex[o]/te:5.1f}%')
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for h in st:
    v = sum(int(r[ix[h]]) for r in data)
    if v > tot * 0.01: print(h, v, f'{100*v/tot:.1f}%')
if len(sys.argv) > 2:
    top = sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:int(sys.argv[2])]
    for r in top:
        reasons = {h[6:]: int(r[ix[h]]) for h in st if int(r[ix[h]]) > 0}
        print(r[ix['# Samples']], r[ix['Source']].strip()[:70], dict(sorted(reasons.items(), key=lambda kv: -kv[1])[:3]))
