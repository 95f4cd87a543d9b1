"""Counts the Blackwell-specific SASS mnemonics per kernel of vmg_b200/libvmgb.so (cuobjdump -sass; CPU only):
UTCHMMA / UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor loads / stores,
UBLKCP = bulk copies, SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0), FFMA2 / FMUL2 = packed FP32x2.
Usage: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'vmg_b200', 'libvmgb.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
pats = ['UTCHMMA.2CTA', 'UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'UCGABAR', 'HMMA', 'FFMA2', 'FMUL2', 'MUFU.TANH', 'LDGSTS', 'REDG', 'ATOMG']
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(.*', '', name).replace('void ', '')
        cur = counts.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        cur['_total'] += 1
        for p in pats:
            if op == p or op.startswith(p + '.') or (p == 'UTCHMMA' and op.startswith('UTCHMMA') and not op.startswith('UTCHMMA.2CTA')):
                cur[p] += 1
                break
arch = subprocess.run(['cuobjdump', '-lelf', so], capture_output=True, text=True).stdout
print('libvmgb.so ELF images:', ', '.join(sorted(set(re.findall(r'sm_\d+a?', arch)))))
print(f'{"kernel":58s} {"instr":>7s} ' + ' '.join(f'{p:>12s}' for p in pats))
for name, c in sorted(counts.items(), key=lambda kv: -kv[1]['_total']):
    if c['_total'] < 50:
        continue
    print(f'{name[:58]:58s} {c["_total"]:7d} ' + ' '.join(f'{c[p]:12d}' for p in pats))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f'{"TOTAL":58s} {tot["_total"]:7d} ' + ' '.join(f'{tot[p]:12d}' for p in pats))
