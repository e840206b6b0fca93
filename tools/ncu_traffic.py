"""Write profiles/r02_ncu_traffic.json: measured DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of one
`ncu --set full` capture) of the kernels bench.py's `roofline.traffic` refers to.

    python tools/ncu_traffic.py <key>=<report.ncu-rep>:<launch index> [...]  > profiles/r02_ncu_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
out = {}
for arg in sys.argv[1:]:
    key, rest = arg.split('=', 1)
    rep, idx = rest.rsplit(':', 1)
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    r = rows[2 + int(idx)]
    tot = 0.0
    for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        tot += float(r[col[m]].replace(',', '')) * SCALE[units[col[m]]]
    out[key] = int(tot)
    out[key + ' (kernel, grid)'] = '%s %s' % (r[col['Kernel Name']][:80], r[col['Grid Size']] if 'Grid Size' in col else '')
print(json.dumps(out, indent=1))
