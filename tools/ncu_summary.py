"""Condense ncu reports (gpurun_out/*/prof_*.ncu-rep) into the small text tables kept under profiles/.

    python tools/ncu_summary.py <report.ncu-rep> [<report2> ...] > profiles/r02_ncu_<what>.txt

One block per profiled launch: duration, DRAM bytes read / written (-> achieved GB/s and the fraction of the measured
copy peak in MEASURED_PEAKS.json), tensor-pipe utilisation, L1/L2 throughput, occupancy, registers, shared memory, grid."""
import csv
import io
import json
import os
import subprocess
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(R, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(R, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}
WANT = [('gpu__time_duration.sum', 'duration'),
        ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram throughput % (ncu peak)'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe % of active cycles'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe % of elapsed cycles'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1/TEX throughput %'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
        ('l1tex__m_xbar2l1tex_read_bytes.sum', 'L2->SM bytes'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
        ('launch__registers_per_thread', 'registers / thread'),
        ('launch__shared_mem_per_block_dynamic', 'dynamic smem / block'),
        ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('sm__cycles_elapsed.avg.per_second', 'SM clock during the capture')]
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1.0, 'msecond': 1e-3,
         'usecond': 1e-6, 'nsecond': 1e-9, 'second': 1.0}
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print('# %s  (ncu --set full --clock-control none; cold-cache single launches: read SHARES and counters, not absolute times)' % os.path.relpath(rep, R))
    for r in rows[2:]:
        name = r[idx['Kernel Name']]
        print('## ' + name[:150])
        vals = {}
        for key, label in WANT:
            if key in idx:
                print('   %-36s %s %s' % (label, r[idx[key]], units[idx[key]]))
                try:
                    vals[key] = float(r[idx[key]].replace(',', '')) * SCALE.get(units[idx[key]], 1.0)
                except ValueError:
                    pass
        if all(k in vals for k in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum')):
            gbs = (vals['dram__bytes_read.sum'] + vals['dram__bytes_write.sum']) / vals['gpu__time_duration.sum'] / 1e9
            print('   %-36s %.0f GB/s = %.2f of the measured copy peak (%.0f GB/s)' % ('DRAM traffic / duration', gbs, gbs / peaks['hbm_gbs'], peaks['hbm_gbs']))
    print()
