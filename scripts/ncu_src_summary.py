"""Opcode / hot-spot summary of an `ncu --page source --csv` dump: python scripts/ncu_src_summary.py file.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
iS, iSm, iI = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


tot_s = sum(num(r[iSm]) for r in data)
tot_i = sum(num(r[iI]) for r in data)
print('kernel', rows[0][1][:80])
print('total samples', tot_s, 'warp instructions', tot_i, 'SASS lines', len(data))
op, ops = collections.Counter(), collections.Counter()
for r in data:
    t = r[iS].split()
    if not t:
        continue
    o = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
    o = o.split('.')[0]
    op[o] += num(r[iI])
    ops[o] += num(r[iSm])
for o, c in op.most_common(28):
    print(f'{o:10s} instr {c / max(tot_i, 1) * 100:5.1f}%  samples {ops[o] / max(tot_s, 1) * 100:5.1f}%')
print('--- top sampled instructions')
for r in sorted(data, key=lambda r: -num(r[iSm]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(r[iSm].rjust(7), r[iI].rjust(9), r[iS][:100])
