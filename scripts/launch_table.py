"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: python scripts/launch_table.py file.csv [last-step]
With `last-step` only the launches of the last optimizer step (between the last two groups of multi_tensor_apply kernels)."""
import collections
import csv
import sys


def read(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    seq = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1000 if u == 'ns' else v * 1000 if u == 'ms' else v
        seq.append((row['Kernel Name'], v, row.get('Grid Size', '')))
    return seq


if __name__ == '__main__':
    seq = read(sys.argv[1])
    if len(sys.argv) > 2 and sys.argv[2] == 'last-step':
        idx = [i for i, (n, v, g) in enumerate(seq) if 'multi_tensor_apply' in n]
        ends = [i for k, i in enumerate(idx) if k + 1 == len(idx) or idx[k + 1] - i > 5]
        seq = seq[ends[-2] + 1:ends[-1] + 1]
    if len(sys.argv) > 3 and sys.argv[3] == 'list':
        for n, v, g in seq:
            if v > 15:
                print(f'{v:9.1f} {g:>14s} {n.split("(")[0][-70:]}')
        sys.exit(0)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v, g in seq:
        k = n.split('(')[0][-80:]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for n, v, g in seq)
    print(f'{len(seq)} launches, {tot / 1000:.2f} ms')
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
        print(f'{v:9.1f} {100 * v / tot:5.1f}% {c:4d}  {k}')
