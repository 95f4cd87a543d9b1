"""Hottest SASS instructions of an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv > f.csv):
python tools/ncu_hot.py f.csv [N] -> top-N instructions by stall samples with their two main stall reasons,
followed by the totals per stall reason."""
import csv
import sys


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main(path, n=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(num(r[idx['# Samples']]) for r in data)
    print('total samples', tot, 'instructions', len(data))
    agg = {h: sum(num(r[idx[h]]) for r in data) for h in stall_cols}
    for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        print(f'  {h:24s} {v:8.0f}  {100 * v / max(tot, 1):5.1f} %')
    top = sorted(enumerate(data), key=lambda ir: -num(ir[1][idx['# Samples']]))[:n]
    for i, r in sorted(top):
        st = sorted(((num(r[idx[h]]), h) for h in stall_cols), reverse=True)[:2]
        print(i, r[idx['Address']][-6:], r[idx['Source']][:80].strip(), int(num(r[idx['# Samples']])), [(h[6:], int(v)) for v, h in st])


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
