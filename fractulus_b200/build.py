"""Builds libfrx_b200.so (hand-written sm_100a CUDA + the C ABI of include/frx.h) in-tree with nvcc.

    python -m fractulus_b200.build [--force] [--verbose]

The shared object lands in fractulus_b200/_build/ (git-ignored, but shipped to the GPU box)."""
import argparse
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'libfrx_b200.so')
SOURCES = ['frx_api.cu', 'frx_weights.cu', 'frx_history.cu', 'frx_solve.cu', 'frx_ode.cu', 'frx_compose.cu', 'frx_bench.cu']
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


def nvcc_path():
    for cand in (os.environ.get('FRX_NVCC'), '/usr/local/cuda/bin/nvcc', shutil.which('nvcc')):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set FRX_NVCC)')


def _fingerprint():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), 'include')):
        for name in sorted(os.listdir(root)):
            if name.endswith(('.cu', '.cuh', '.h', '.inc')):
                with open(os.path.join(root, name), 'rb') as f:
                    h.update(name.encode())
                    h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, 'fingerprint.txt')
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == fp:
        return LIB
    nvcc = nvcc_path()
    common = [nvcc, '-O3', '-std=c++17', '-lineinfo', '-ccbin', '/usr/bin/g++', '-Xcompiler', '-fPIC',
              '-Xcompiler', '-ffp-contract=off'] + ARCH
    if verbose:
        common += ['-Xptxas', '-v']
    objs, procs = [], []
    for src in SOURCES:   # translation units compile in parallel
        obj = os.path.join(OUT_DIR, src[:-3] + '.o')
        cmd = common + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = []
    for src, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode:
            sys.stderr.write(out)
        if pr.returncode:
            failed.append(src)
    if failed:
        raise RuntimeError('nvcc failed on ' + ', '.join(failed))
    cmd = [nvcc, '-shared', '-ccbin', '/usr/bin/g++', '-o', LIB] + objs + ARCH
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('link failed')
    with open(stamp, 'w') as f:
        f.write(fp)
    return LIB


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    a = ap.parse_args()
    print(build(a.force, a.verbose))
