"""Summarise an .ncu-rep here (no GPU needed): headline metrics + the instructions with the most stall samples.
Usage: python tools/ncu_stalls.py gpurun_out/x.ncu-rep [top_n] [kernel_index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'lts__t_bytes.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second']
for i, h in enumerate(hdr):
    if h in want:
        print(f'{h:75s} {rows[1][i]:>10s} ' + ' | '.join(r[i] for r in rows[2:]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# sections: a "Kernel Name" row, then a header row starting with "Address", then one row per SASS instruction
secs = []
for i, r in enumerate(rows):
    if r and r[0] == 'Kernel Name':
        secs.append([r[1], None, None])
    elif r and r[0] == 'Address' and secs and secs[-1][1] is None:
        secs[-1][1] = i
for k, sc in enumerate(secs):
    sc[2] = (secs[k + 1][1] - 2) if k + 1 < len(secs) and secs[k + 1][1] else len(rows)
secs = [sc for sc in secs if sc[1] is not None]
if len(sys.argv) <= 3:
    for k, sc in enumerate(secs):
        print(f'section {k}: {sc[0][:100]}')
name, h0, end = secs[kidx]
h = rows[h0]
body = [r for r in rows[h0 + 1:end] if len(r) == len(h)]
ci = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(int(r[ci['# Samples']] or 0) for r in body) or 1
print(f'\nsection {kidx}: {name[:90]}: {tot} stall samples over {len(body)} instructions')
agg = {n: sum(int(r[ci[n]] or 0) for r in body) for n in stalls}
print('by reason:', ', '.join(f'{k[6:]} {100 * v / tot:.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(body, key=lambda r: -int(r[ci['# Samples']] or 0))[:top_n]:
    s = int(r[ci['# Samples']])
    st = sorted(((n[6:], int(r[ci[n]] or 0)) for n in stalls), key=lambda kv: -kv[1])[:2]
    print(f"{100 * s / tot:5.1f}% {r[ci['Address']][-5:]} ex={r[ci['Instructions Executed']]:>9s} {r[ci['Source']][:72]:72s} {[x for x in st if x[1]]}")
