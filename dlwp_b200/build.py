"""
Build libdlwp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m dlwp_b200.build [--force] [--verbose]

The shared library depends only on the CUDA runtime (linked statically) -- no torch, no libcuda at link time (the one
driver call, cuTensorMapEncodeTiled, is resolved through cudaGetDriverEntryPoint).
"""

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(CSRC, 'libdlwp_b200.so')
SOURCES = ['conv.cu', 'conv_tc.cu', 'conv_sw_net_a.cu', 'conv_sw_net_b.cu', 'conv_sw_net_basic.cu', 'conv_sw_bf16.cu', 'conv_fused.cu',
           'elementwise.cu', 'plan.cu', 'train.cu']
HEADERS = ['internal.h', 'conv_tc.h', 'conv_sw.cuh', os.path.join('..', '..', 'include', 'dlwp_b200.h')]
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    flags = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--use_fast_math=false'] + ARCH
    flags = [f for f in flags if f != '--use_fast_math=false']
    if verbose:
        flags += ['-Xptxas', '-v']
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [_nvcc()] + flags + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-cudart', 'static'] + ARCH
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
