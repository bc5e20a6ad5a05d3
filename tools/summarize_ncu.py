#!/usr/bin/env python
"""Table of the key metrics of an `ncu --set full` report (per launch; `tensor busy %` = hmma sub-pipe active cycles / 4 / elapsed SM cycles).  usage: summarize_ncu.py <report.ncu-rep> > table.md  (and writes the raw
CSV of the selected columns next to it when a second argument is given)."""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
idx = {n: i for i, n in enumerate(h)}
COLS = [('time us', 'gpu__time_duration.sum'), ('dram rd MB', 'dram__bytes_read.sum'), ('dram wr MB', 'dram__bytes_write.sum'),
        ('dram %', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), ('L2 %', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
        ('L2->SM GB', 'lts__t_sectors_srcunit_tex_op_read.sum'),
        ('tensor busy %', '__tensor_busy__'),
        ('tc smem wavefronts %', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
        ('issue %', 'sm__inst_issued.avg.pct_of_peak_sustained_active'), ('warps active %', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
        ('regs', 'launch__registers_per_thread'), ('smem KB', 'launch__shared_mem_per_block_dynamic')]
def find(name):
    if name in idx: return idx[name]
    for k in idx:
        if k.endswith(name): return idx[k]
    return None
def val(row, name):
    if name == '__tensor_busy__':
        # UTCHMMA keeps the four tensor sub-pipes of an SM busy: busy fraction = hmma sub-pipe active cycles / 4 / elapsed SM cycles
        # (sm__pipe_tensor_cycles_active_realtime.pct, which round 1 printed, under-reports tcgen05 work by ~3x)
        a, b = find('sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg'), find('sm__cycles_elapsed.avg')
        if a is None or b is None or row[a] in ('', 'no data', 'n/a') or row[b] in ('', 'no data', 'n/a'):
            return ''
        return f"{100.0 * float(row[a].replace(',', '')) / 4.0 / float(row[b].replace(',', '')):.1f}"
    i = find(name)
    if i is None or row[i] in ('', 'no data', 'n/a'): return ''
    v = float(row[i].replace(',', ''))
    u = units[i]
    if name.startswith('dram__bytes'):
        v *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(u, 1)
    if name == 'gpu__time_duration.sum':
        v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1, 'msecond': 1e3}.get(u, 1)
    if name.startswith('lts__t_sectors'):
        v = v * 32 / 1e9
    if name == 'launch__shared_mem_per_block_dynamic':
        v *= {'byte': 1 / 1024, 'Kbyte': 1, 'Mbyte': 1024}.get(u, 1)
    return f'{v:.1f}' if abs(v) < 1e4 else f'{v:.0f}'
print('| kernel | grid | ' + ' | '.join(c for c, _ in COLS) + ' |')
print('|---|---|' + '---|' * len(COLS))
out_rows = []
for row in rows[2:]:
    name = re.sub(r'^(void )?(b200gan::)?(\(anonymous namespace\)::|<unnamed>::)?', '', row[idx['Kernel Name']]).split('(')[0][:56]
    vals = [val(row, n) for _, n in COLS]
    print(f"| `{name}` | {row[idx['Grid Size']]} | " + ' | '.join(vals) + ' |')
    out_rows.append([name, row[idx['Grid Size']]] + vals)
if len(sys.argv) > 2:
    with open(sys.argv[2], 'w', newline='') as f:
        w = csv.writer(f); w.writerow(['kernel', 'grid'] + [c for c, _ in COLS]); w.writerows(out_rows)
