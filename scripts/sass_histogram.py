"""SASS opcode histogram of the built library: python scripts/sass_histogram.py > profiles/r02_sass_histogram.md
Runs `cuobjdump -sass` on tfep_b200/lib/libtfep_b200.so (no GPU needed) and counts the opcodes per kernel."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'tfep_b200', 'lib', 'libtfep_b200.so')
BLACKWELL = ('UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'UTCATOMSWS', 'MUFU', 'ELECT', 'USETMAXREG', 'FFMA2', 'FADD2', 'FMUL2')


def kernels():
    out = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True, check=True).stdout
    name, ops = None, None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            if name:
                yield name, ops
            name, ops = m.group(1), collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)', line)
        if m and name:
            ops[m.group(1)] += 1
            if m.group(1) in ('MUFU', 'SYNCS'):
                ops[m.group(1) + m.group(2)] += 0   # keep the base count only; variants listed separately below
                ops['~' + m.group(1) + m.group(2)] += 1
    if name:
        yield name, ops


def demangle(n):
    return subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip().split('(')[0]


if __name__ == '__main__':
    ks = list(kernels())
    total = collections.Counter()
    for _, ops in ks:
        total.update({k: v for k, v in ops.items() if not k.startswith('~')})
    print('# r02 SASS opcode histogram of `tfep_b200/lib/libtfep_b200.so` (sm_100a)\n')
    print('`python scripts/sass_histogram.py` = `cuobjdump -sass` of the in-tree library, opcodes counted per kernel '
          '(static counts of the compiled code, not executed instructions).\n')
    print(f'{len(ks)} kernels, {sum(total.values())} instructions.  Blackwell-specific opcodes over the whole library:\n')
    print('| opcode | count | meaning |\n|---|---|---|')
    meaning = {'UTCHMMA': 'tcgen05.mma (kind::f16)', 'UTCBAR': 'tcgen05.commit -> mbarrier', 'LDTM': 'tcgen05.ld (TMEM -> registers)',
               'STTM': 'tcgen05.st (registers -> TMEM)', 'UBLKCP': 'cp.async.bulk (TMA engine, linear)', 'SYNCS': 'mbarrier operations',
               'MUFU': 'special-function unit (ex2 / lg2 / rcp ...)', 'ELECT': 'elect.sync', 'USETMAXREG': 'setmaxnreg',
               'UTCATOMSWS': 'tcgen05.alloc / dealloc', 'FFMA2': 'fma.rn.f32x2 (packed fp32)', 'FADD2': 'add.rn.f32x2', 'FMUL2': 'mul.rn.f32x2', 'UTMALDG': 'cp.async.bulk.tensor load', 'UTMASTG': 'cp.async.bulk.tensor store'}
    for op in BLACKWELL:
        if total[op]:
            print(f'| `{op}` | {total[op]} | {meaning.get(op, "")} |')
    print('\n## Per kernel (kernels that use the tensor cores or bulk copies)\n')
    print('| kernel | instructions | UTCHMMA | LDTM | STTM | UBLKCP | UTCBAR | SYNCS | MUFU | FFMA+FMUL+FADD | top opcodes |\n|---|---|---|---|---|---|---|---|---|---|---|')
    for name, ops in sorted(ks, key=lambda kv: -sum(v for k, v in kv[1].items() if not k.startswith('~'))):
        base = {k: v for k, v in ops.items() if not k.startswith('~')}
        if not (base.get('UTCHMMA') or base.get('UBLKCP')):
            continue
        n = sum(base.values())
        top = ', '.join(f'{k} {v}' for k, v in sorted(base.items(), key=lambda kv: -kv[1])[:8])
        fp = base.get('FFMA', 0) + base.get('FMUL', 0) + base.get('FADD', 0)
        print(f'| `{demangle(name)[-90:]}` | {n} | {base.get("UTCHMMA", 0)} | {base.get("LDTM", 0)} | {base.get("STTM", 0)} | '
              f'{base.get("UBLKCP", 0)} | {base.get("UTCBAR", 0)} | {base.get("SYNCS", 0)} | {base.get("MUFU", 0)} | {fp} | {top} |')
    # the headline kernel in full
    for name, ops in ks:
        d = demangle(name)
        if 'maf_spline_fwd_kernel<false, false>' in d:
            base = {k: v for k, v in ops.items() if not k.startswith('~')}
            print(f'\n## Headline kernel `{d}`: all opcodes\n')
            print(', '.join(f'`{k}` {v}' for k, v in sorted(base.items(), key=lambda kv: -kv[1])))
            print('\nMUFU / SYNCS variants: ' + ', '.join(f'`{k[1:]}` {v}' for k, v in sorted(ops.items()) if k.startswith('~')))
