"""Build the C-ABI CUDA library in-tree: tfep_b200/lib/libtfep_b200.so (sm_100a only).

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting .so is
git-ignored but travels with the repository snapshot to the GPU box.
"""

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIBPATH = os.path.join(LIBDIR, 'libtfep_b200.so')
OBJDIR = os.path.join(HERE, 'build')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

SOURCES = ['core.cu', 'gemm_simt.cu', 'transformers.cu', 'analysis.cu', 'maf_fused_sm100.cu', 'maf_fused_inv_sm100.cu', 'maf_inverse.cu', 'tc_gemm_sm100.cu', 'frames.cu', 'loss.cu', 'mt19937_jump.cu', 'weights.cu']

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC',
    '-Xptxas', '-v',
    '-I', INCLUDE,
] + os.environ.get('TFEPB_EXTRA_NVCC_FLAGS', '').split()

# Per-file flags.  tc_gemm_sm100.cu: the transformer epilogues of the bf16 tensor-core path run on 16 epilogue warps per SM
# next to the MMA pipeline, where the IEEE division / sqrt / log sequences of the exact kernels are latency bound; their
# results carry bf16 operand rounding anyway (stated tolerance in DESIGN.md), so this file alone uses the approximate
# MUFU forms (rcp / rsq / lg2 / ex2, ~1e-6 relative).  The exact fp32 / fp64 kernels are compiled without it.
FILE_FLAGS = {'tc_gemm_sm100.cu': ['--use_fast_math']}


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; the tfep_b200 CUDA library cannot be built')
    return exe


def _deps(src):
    deps = [src, os.path.join(INCLUDE, 'tfep_b200.h')]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    return deps


def _signature(paths, extra=()):
    extra = list(extra)
    h = hashlib.sha1()
    h.update(' '.join(NVCC_FLAGS + extra).encode())
    for p in sorted(paths):
        with open(p, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def _compile_one(name, verbose):
    src = os.path.join(CSRC, name)
    obj = os.path.join(OBJDIR, name.replace('.cu', '.o'))
    sig_file = obj + '.sig'
    extra = FILE_FLAGS.get(name, [])
    sig = _signature(_deps(src), extra)
    if os.path.exists(obj) and os.path.exists(sig_file) and open(sig_file).read() == sig:
        return obj, ''
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ['-c', src, '-o', obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    if res.returncode != 0:
        raise RuntimeError(f'nvcc failed for {name}:\n{log}')
    with open(sig_file, 'w') as f:
        f.write(sig)
    with open(obj + '.log', 'w') as f:
        f.write(log)
    if verbose:
        print(log)
    return obj, log


def build(verbose=False, force=False):
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
    os.makedirs(OBJDIR, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = [o for o, _ in ex.map(lambda n: _compile_one(n, verbose), sources)]
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIBPATH) or os.path.getmtime(LIBPATH) < newest:
        cmd = [_nvcc(), '-shared', '-o', LIBPATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcuda']
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError('link failed:\n' + res.stdout + res.stderr)
    return LIBPATH


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
