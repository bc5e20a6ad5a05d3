#!/usr/bin/env python
"""SASS evidence of the Blackwell-native paths: per kernel of libb200gan.so, the number of tcgen05 MMAs (UTCHMMA, .2CTA counted apart),
TMEM loads (LDTM), TMA loads / stores (UTMALDG / UTMASTG), tcgen05 commits (UTCBAR) and legacy warp-level MMAs (HMMA) in the sm_100a
code.  usage: sass_histogram.py [path/to/libb200gan.so] > profiles/rNN_sass_histogram.md   (no GPU needed: cuobjdump -sass)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gan-enhanced-pneumonia-classifier_b200', 'libb200gan.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
OPS = ['UTCHMMA.2CTA', 'UTCHMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UTCBAR', 'HMMA', 'LDGSTS']
rows, cur, k = collections.OrderedDict(), None, 0
for line in sass.split('\n'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = names[k] if k < len(names) else m.group(1)
        k += 1
        cur = re.sub(r'\((int|bool)\)', '', cur)
        cur = re.sub(r'^(void )?(b200gan::)?(\(anonymous namespace\)::|<unnamed>::)?', '', cur).split('(')[0]
        rows.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r'^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if not m:
        continue
    op = m.group(1)
    if op.startswith('UTCHMMA'):
        rows[cur]['UTCHMMA.2CTA' if '.2CTA' in op else 'UTCHMMA'] += 1
    elif op.startswith('HMMA'):
        rows[cur]['HMMA'] += 1
    else:
        for o in ('LDTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'LDGSTS'):
            if op.startswith(o):
                rows[cur][o] += 1
print(f'SASS op histogram of `{os.path.relpath(lib, ROOT)}` (sm_100a; `cuobjdump -sass`), kernels with at least one tensor / TMA instruction\n')
print('| kernel | ' + ' | '.join(OPS) + ' |')
print('|---|' + '---:|' * len(OPS))
tot = collections.Counter()
for name, c in rows.items():
    if not any(c[o] for o in OPS if o != 'LDGSTS'):
        continue
    tot.update(c)
    print(f'| `{name[:90]}` | ' + ' | '.join(str(c[o]) if c[o] else '' for o in OPS) + ' |')
print('| **total** | ' + ' | '.join(str(tot[o]) for o in OPS) + ' |')
