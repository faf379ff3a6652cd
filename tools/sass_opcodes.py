"""Opcode histogram per kernel of libvlnimagine.so (cuobjdump -sass): which kernels carry the Blackwell tensor-core / TMA opcodes.
UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add,
UTCBAR = tcgen05.commit, HMMA = mma.sync, LDGSTS = cp.async.   python tools/sass_opcodes.py > profiles/rNN_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'vln-imagine_b200', 'libvlnimagine.so')
OPS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'HMMA', 'LDGSTS', 'LDSM', 'MUFU']
out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
demangle = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', out)), capture_output=True, text=True).stdout.split('\n')
names = iter(demangle)
counts, cur, total = collections.OrderedDict(), None, collections.Counter()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = next(names, m.group(1))
        cur = re.sub(r'\(anonymous namespace\)::', '', cur)
        cur = re.sub(r'\((?!bool\)|int\)).*$', '', cur).replace('void ', '').replace('<unnamed>::', '').replace('(bool)', '').replace('(int)', '')
        counts.setdefault(cur, collections.Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur is not None:
        op = m.group(1)
        total[cur] += 1
        for o in OPS:
            if op.startswith(o):
                counts[cur][o] += 1
print('# SASS opcode histogram of libvlnimagine.so (sm_100a), per kernel\n')
print('| kernel | instructions | ' + ' | '.join(OPS) + ' |')
print('|---|---|' + '---|' * len(OPS))
for k, c in counts.items():
    if total[k] == 0:
        continue
    print('| `%s` | %d | ' % (k[:90], total[k]) + ' | '.join(str(c[o]) if c[o] else '' for o in OPS) + ' |')
