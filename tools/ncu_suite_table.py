"""Turns the raw-page CSV of the ncu_suite capture into the per-kernel table of profiles/r2_suite_ncu.md:
duration, DRAM bytes read/written, achieved DRAM GB/s against the measured copy peak, tensor-pipe and issue activity.
Usage: python tools/ncu_suite_table.py gpurun_out/r2_suite_raw.csv gpurun_out/r2m_plain.log > profiles/r2_suite_ncu.md"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
names = [l.split()[1] for l in open(sys.argv[2]) if l.startswith('ran ')] if len(sys.argv) > 2 else []
h, units = rows[0], rows[1]
rows = rows[:2] + [r for r in rows[2:] if 'finalize' not in r[h.index('Kernel Name')]]      # helper kernels are not cases
ci = {n: i for i, n in enumerate(h)}
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbs', 6446.3) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6446.3


def val(r, k, unit_scale=None):
    v = float(r[ci[k]].replace(',', '')) if r[ci[k]] not in ('', 'n/a') else float('nan')
    u = units[ci[k]]
    if unit_scale:
        v *= unit_scale.get(u, 1.0)
    return v


BY = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
TM = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}
print('| # | case | kernel | us | DRAM MB (read+write) | - | DRAM GB/s | of measured HBM peak (%.0f GB/s) | tensor pipe %% | issue active %% | regs | grid x block |' % peak)
print('|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---|')
k = 0
for r in rows[2:]:
    kn = r[ci['Kernel Name']].split('(')[0].replace('void ', '')
    if kn.startswith('morph_block16') and k == 0 and names and not names[0].startswith('morph'):
        pass
    us = val(r, 'gpu__time_duration.sum', TM)
    if 'dram__bytes_read.sum' in ci:
        rd, wr = val(r, 'dram__bytes_read.sum', BY), val(r, 'dram__bytes_write.sum', BY)
        gbs = (rd + wr) / us / 1e3
    else:       # section-based capture: DRAM bandwidth only (read + write); bytes = bandwidth x duration
        gbs = val(r, 'dram__bytes.sum.per_second', {'byte/s': 1e-9, 'Kbyte/s': 1e-6, 'Mbyte/s': 1e-3, 'Gbyte/s': 1.0, 'Tbyte/s': 1e3})
        rd, wr = gbs * us * 1e3, 0.0
    tp = val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed') if 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed' in ci else float('nan')
    if tp != tp and 'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active' in ci:
        tp = val(r, 'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active')
    ia = val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active') if 'smsp__issue_active.avg.pct_of_peak_sustained_active' in ci else float('nan')
    idx = k
    k += 1
    case = names[idx - (len(rows) - 2 - len(names))] if names and idx >= len(rows) - 2 - len(names) else '(setup)'
    print(f"| {idx} | {case} | `{kn[:44]}` | {us:.1f} | {(rd + wr) / 1e6:.1f} | | {gbs:.0f} | {gbs / peak:.2f} | {tp:.1f} | {ia:.1f} | {r[ci['launch__registers_per_thread']]} | {r[ci['launch__grid_size']]} x {r[ci['launch__block_size']]} |")
