"""Per-kernel SASS opcode counts of the in-tree shared object: the mnemonics that prove a Blackwell-native kernel
(B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = legacy mma.sync).  Writes profiles/<tag>_sass_opcounts.txt.

    python tools/sass_opcounts.py [tag=r02]"""
import collections
import os
import re
import subprocess
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(R, 'efficientdet.pytorch_b200', 'csrc', 'libeffdet_b200.so')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
KEYS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UBLKCP', 'SYNCS', 'LDGSTS', 'HMMA', 'LDG', 'STG', 'RED', 'ATOMG',
        'LDS', 'STS', 'MUFU', 'FFMA', 'BAR', 'LD', 'ST']      # LD / ST = generic-address loads / stores (exact match)
out = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True, check=True).stdout
kern, counts = None, {}
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r'^void ', '', kern)
        kern = re.sub(r'\(.*$', '', kern)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and kern:
        op = m.group(1)
        for k in KEYS:
            if (op == k) if k in ('LD', 'ST') else op.startswith(k):
                counts[kern][k] += 1
                break
rows = sorted(counts.items(), key=lambda kv: (-kv[1]['UTCHMMA'], -kv[1]['UTMALDG'], kv[0]))
path = os.path.join(R, 'profiles', '%s_sass_opcounts.txt' % tag)
with open(path, 'w') as f:
    f.write('# cuobjdump -sass %s | opcode prefixes per kernel (tools/sass_opcounts.py)\n' % os.path.relpath(SO, R))
    f.write('%-78s ' % 'kernel' + ' '.join('%7s' % k for k in KEYS) + '\n')
    for kname, c in rows:
        f.write('%-78s ' % kname[:78] + ' '.join('%7d' % c[k] for k in KEYS) + '\n')
    tot = collections.Counter()
    for _, c in rows:
        tot.update(c)
    f.write('%-78s ' % 'TOTAL' + ' '.join('%7d' % tot[k] for k in KEYS) + '\n')
print(open(path).read()[:3000])
