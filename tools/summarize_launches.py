"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, total and
mean device time, share of the profiled region.  Usage: python tools/summarize_launches.py launches.csv [out.md]"""
import collections
import csv
import sys


def main(path, out=None):
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(unit, 1)
        rows.append((r['Kernel Name'].split('(')[0], ns))
    agg = collections.OrderedDict()
    for k, ns in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values()) or 1.0
    lines = ['| kernel | launches | total ms | mean us | share |', '|---|---:|---:|---:|---:|']
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f'| `{k}` | {n} | {ns / 1e6:.2f} | {ns / n / 1e3:.1f} | {100 * ns / tot:.1f} % |')
    lines.append(f'| **total** | {len(rows)} | {tot / 1e6:.2f} | | 100 % |')
    text = '\n'.join(lines)
    print(text)
    if out:
        open(out, 'w').write(text + '\n')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
