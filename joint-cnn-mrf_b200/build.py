"""Builds libjcm.so (the sm_100a kernels + C ABI) in-tree with nvcc.  No GPU needed: nvcc cross-compiles.

    python joint-cnn-mrf_b200/build.py [--force] [--verbose]
"""
import argparse
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'jcm', 'csrc')
OUT_LIB = os.path.join(HERE, 'jcm', 'libjcm.so')
OUT_EXP = os.path.join(HERE, 'jcm', 'libjcm_exp.so')
STAMP_LIB = os.path.join(HERE, 'jcm', '.libjcm.stamp')
SOURCES = ['core.cu', 'tiling.cu', 'prep.cu', 'glue.cu', 'conv_tcgen05.cu', 'spatial_model.cu', 'backward.cu', 'taps.cu', 'optim.cu', 'augment.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '--expt-relaxed-constexpr',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-Wall', '-Xcompiler', '-Wno-unused-function']


def _digest(srcs):
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith(('.cu', '.cuh', '.h')):
            h.update(f.encode())
            h.update(open(os.path.join(CSRC, f), 'rb').read())
    h.update(' '.join(NVCC_FLAGS + srcs).encode())
    return h.hexdigest()


def build(force=False, verbose=False, experiments=False):
    """experiments=True builds libjcm_exp.so with -DJCM_EXPERIMENTS (environment-variable measurement switches and the CTA-pair
    experiment compiled in); it is loaded only by tests/gpu_diag.py, never by the jcm package."""
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    flags = NVCC_FLAGS + (['-DJCM_EXPERIMENTS=1'] if experiments else [])
    OUT = OUT_EXP if experiments else OUT_LIB
    STAMP = OUT + '.stamp' if experiments else STAMP_LIB
    dig = _digest(srcs + flags)
    if not force and os.path.exists(OUT) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    bdir = os.path.join(HERE, 'build_exp' if experiments else 'build')
    os.makedirs(bdir, exist_ok=True)
    for s in srcs:
        o = os.path.join(bdir, s.replace('.cu', '.o'))
        cmd = [nvcc] + flags + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, s), '-o', o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write('---- %s\n%s\n' % (s, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [nvcc, '-shared', '-o', OUT] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart_static', '-ldl', '-lrt',
                                                  '-lpthread']
    subprocess.check_call(cmd)
    with open(STAMP, 'w') as f:
        f.write(dig)
    return OUT


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    ap.add_argument('--experiments', action='store_true', help='build libjcm_exp.so (-DJCM_EXPERIMENTS) for tests/gpu_diag.py')
    a = ap.parse_args()
    print(build(a.force, a.verbose, a.experiments))
