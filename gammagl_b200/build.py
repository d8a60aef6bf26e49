"""In-tree build of libgglb200.so (hand-written sm_100a CUDA behind the C ABI of include/gglb200.h).

    python -m gammagl_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; each translation unit is compiled in parallel into build/obj and
linked into gammagl_b200/libgglb200.so (git-ignored, but it travels to the GPU box with gpurun).
The library has no dependency on torch / libtorch -- only the CUDA runtime (linked statically).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(ROOT, 'build', 'obj')
LIB = os.path.join(PKG, 'libgglb200.so')
STAMP = os.path.join(ROOT, 'build', 'libgglb200.stamp')

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
CFLAGS = ['-O3', '-std=c++17', '-lineinfo', '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC',
          '-Xcompiler', '-fvisibility=hidden', '-Xcompiler', '-Wno-unused-function']


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ['../../include/gglb200.h']:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, 'rb').read())
    h.update(' '.join(ARCH + CFLAGS).encode())
    return h.hexdigest()


def kernel_digest(files=('spmm_impl.cuh', 'common.cuh', 'spmm_f32.cu')):
    """Digest of the sources ONE kernel family is compiled from (default: spmm_ring_kernel<float, ...>) plus the tuning constants of
    the public header and the compiler flags.  An ncu capture stamped with it (tools/capture_round.sh) stays valid while exactly
    these sources are unchanged; bench.py refuses a capture whose stamp differs (roofline.traffic)."""
    h = hashlib.sha256()
    for f in files:
        h.update(f.encode())
        h.update(open(os.path.join(CSRC, f), 'rb').read())
    for line in open(os.path.join(ROOT, 'include', 'gglb200.h')):
        if line.startswith('#define GGLB200_') and 'ABI_VERSION' not in line:
            h.update(line.encode())
    h.update(' '.join(ARCH + CFLAGS).encode())
    return h.hexdigest()


def up_to_date():
    return os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == _digest()


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):
            return LIB  # GPU box without a toolkit: use the shipped binary
        raise RuntimeError('nvcc not found and no prebuilt libgglb200.so')
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + '.o')
        cmd = [NVCC] + ARCH + CFLAGS + (['-Xptxas', '-v'] if verbose else []) + \
              ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    tmp = LIB + '.tmp.%d' % os.getpid()          # link next to the target, then rename: a reader never sees a half-written library
    cmd = [NVCC] + ARCH + ['-shared', '-o', tmp] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    os.replace(tmp, LIB)
    with open(STAMP, 'w') as f:
        f.write(_digest())
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
