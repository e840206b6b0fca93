"""Source-level stall summary of an `ncu --set full --import-source on` report: per captured launch the warp-stall sample
totals by reason and the instructions that collected the most samples.

    python tools/ncu_stalls.py <report.ncu-rep> [top=20] > profiles/r02_ncu_<what>_stalls.txt
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-units', 'base'], capture_output=True, text=True).stdout
secs, cur = [], None
for r in csv.reader(io.StringIO(txt)):
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        secs.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
print('# %s : warp-stall samples per launch (ncu source page)' % rep)
for n, sec in enumerate(secs):
    if not sec['rows']:
        continue
    hdr = sec['rows'][0]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in sec['rows'][1:] if len(r) == len(hdr)]
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = {s: sum(int(r[idx[s]] or 0) for r in data) for s in stalls}
    total = sum(int(r[idx['# Samples']] or 0) for r in data)
    inst = sum(int(r[idx['Instructions Executed']] or 0) for r in data)
    print('## launch %d: %s' % (n, sec['name'][:110]))
    print('   samples %d, warp instructions executed %d' % (total, inst))
    print('   ' + ', '.join('%s %.1f%%' % (s.replace('stall_', ''), 100.0 * v / max(total, 1)) for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:7]))
    for r in sorted(data, key=lambda r: -int(r[idx['# Samples']] or 0))[:top_n]:
        st = sorted(((int(r[idx[s]] or 0), s.replace('stall_', '')) for s in stalls), reverse=True)[:2]
        print('   %6s smp %9s exec  %-74s %s' % (r[idx['# Samples']], r[idx['Instructions Executed']], r[idx['Source']].strip()[:74],
                                                ' '.join('%s:%d' % (s, v) for v, s in st)))
