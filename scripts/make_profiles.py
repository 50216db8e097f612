"""Turn .ncu-rep captures / launch lists in gpurun_out/ into the committed summaries under profiles/.

usage: python scripts/make_profiles.py details <rep> <out.md> "<title>" "<command / description>"
       python scripts/make_profiles.py launches <csv> <out.md> "<title>" "<footer>"
"""
import collections, csv, subprocess, sys


def details(rep, out, title, desc):
    raw = subprocess.run(f"ncu -i {rep} --page details --csv", shell=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {title}", "", desc, "", "| section | metric | unit | value |", "|---|---|---|---|"]
    for r in rows[1:]:
        if len(r) <= ix['Metric Value']:
            continue
        lines.append(f"| {r[ix['Section Name']]} | {r[ix['Metric Name']]} | {r[ix['Metric Unit']]} | {r[ix['Metric Value']]} |")
    summ = subprocess.run(f"python scripts/ncu_summary.py {rep} x", shell=True, capture_output=True, text=True).stdout
    lines += ["", "## raw metrics, stall reasons per issued instruction, stall samples by opcode", "", "```", summ.rstrip(), "```"]
    open(out, 'w').write("\n".join(lines) + "\n")


def launches(path, out, title, footer):
    rows = list(csv.reader(open(path)))
    start = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[start]; ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0.0, 0])
    for r in rows[start + 2:]:
        if len(r) <= ix['Metric Value']:
            continue
        try:
            v = float(r[ix['Metric Value']].replace(',', ''))
        except ValueError:
            continue
        unit = r[ix['Metric Unit']]
        if unit in ('ns', 'nsecond'):
            v /= 1000
        elif unit in ('ms', 'msecond'):
            v *= 1000
        k = r[ix['Kernel Name']][:100]
        agg[k][0] += v; agg[k][1] += 1
    tot = sum(v[0] for v in agg.values())
    L = [f"# {title}", "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv` after the same command exited 0 without "
         "ncu.  First 300 launches of the process; per-launch times are cold-cache and serialised: compare SHARES.", "",
         "| share | total us | launches | avg us | kernel |", "|---|---|---|---|---|"]
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
        L.append(f"| {100 * t / tot:.1f}% | {t:.1f} | {n} | {t / n:.1f} | `{k}` |")
    L += ["", footer]
    open(out, 'w').write("\n".join(L) + "\n")


if __name__ == '__main__':
    {'details': details, 'launches': launches}[sys.argv[1]](*sys.argv[2:6])
