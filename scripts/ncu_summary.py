"""Print the key numbers of an .ncu-rep (first kernel) and the hottest source lines by stall samples."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals))
keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active']
for k in keys:
    if k in d:
        print(f'{k:75s} {d[k]} {units[hdr.index(k)]}')
print('--- stall reasons per issue ---')
for h, v in zip(hdr, vals):
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
        try:
            if float(v) > 0.05:
                print(f'   {h[34:-23]:28s} {float(v):.2f}')
        except ValueError:
            pass
if len(sys.argv) > 2:
    src = subprocess.run(f"ncu -i {rep} --page source --csv --print-source sass", shell=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[0]
    try:
        ci = hdr.index('# Samples') if '# Samples' in hdr else [i for i, h in enumerate(hdr) if 'Sampl' in h][0]
    except IndexError:
        print(hdr); sys.exit()
    si = hdr.index('Source')
    tot = sum(float(r[ci] or 0) for r in rows[1:] if len(r) > ci)
    agg = collections.Counter()
    for r in rows[1:]:
        if len(r) > ci:
            op = r[si].split()[0] if r[si].split() else ''
            if op.startswith('@'):
                op = r[si].split()[1]
            agg[op.split('.')[0]] += float(r[ci] or 0)
    print('--- samples by opcode ---')
    for k, v in agg.most_common(18):
        print(f'   {k:14s} {100*v/tot:5.1f}%')
