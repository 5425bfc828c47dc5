#!/usr/bin/env python
"""Instruction / sample shares over coarse SASS index ranges: tools/ncu_ranges.py rep b0,b1,b2,... [kernel-regex]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
bounds = [int(x) for x in sys.argv[2].split(',')]
kre = sys.argv[3] if len(sys.argv) > 3 else None
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + (['--kernel-name', 'regex:' + kre] if kre else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hi[0]]
d = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
ci = {k: i for i, k in enumerate(h)}
def f(r, k):
    try: return float(r[ci[k]])
    except Exception: return 0.0
tot = sum(f(r, 'Instructions Executed') for r in d); ts = sum(f(r, '# Samples') for r in d)
bounds = bounds + [len(d)]
stall_keys = [k for k in ci if k.startswith('stall_') and 'Not Issued' not in k]
for a, b in zip(bounds[:-1], bounds[1:]):
    seg = d[a:b]
    e = sum(f(r, 'Instructions Executed') for r in seg); s = sum(f(r, '# Samples') for r in seg)
    th = sum(f(r, 'Thread Instructions Executed') for r in seg)
    st = sorted(((sum(f(r, k) for r in seg), k[6:]) for k in stall_keys), reverse=True)[:5]
    print('sass[%5d..%5d) instr %5.1f%%  samples %5.1f%%  lanes %4.1f  cyc/instr-rel %.2f   %s' % (
        a, b, 100 * e / tot, 100 * s / ts, th / max(e, 1), (s / ts) / max(e / tot, 1e-9),
        ' '.join('%s=%.0f%%' % (k, 100 * v / max(s, 1)) for v, k in st)))
