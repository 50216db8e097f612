"""Per-CUDA-line instruction / stall-sample histogram from
   ncu -i rep --page source --csv --print-source sass,cuda --kernel-id :::N > file.csv
usage: python scripts/ncu_line_summary.py file.csv [n_lines]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
cur, hdr, out = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] in ('File Path', 'File Name'):
        cur, hdr = r[1], None
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] != '':
        try:
            i, s = hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
            out.append((cur.split('/')[-1], int(r[0]), r[1].strip()[:100], int(r[i]), int(r[s])))
        except ValueError:
            pass
tot, ts = sum(o[3] for o in out), sum(o[4] for o in out)
print('total warp instructions', tot, 'samples', ts)
for o in sorted(out, key=lambda o: -o[3])[:int(sys.argv[2]) if len(sys.argv) > 2 else 50]:
    print(f'{o[0][:16]:16s} {o[1]:5d} {100 * o[3] / tot:5.1f}% s{100 * o[4] / max(ts, 1):5.1f}%  {o[2]}')
