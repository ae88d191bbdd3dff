"""Builds libvodagg.so (hand-written sm_100a CUDA kernels + C ABI) in-tree with nvcc.

    python -m lowlightenvironmentvideoobjectdetection_b200.build [--force]

nvcc cross-compiles without a GPU; the .so lands next to this file so it travels
with the repo snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB_PATH = os.path.join(_HERE, 'libvodagg.so')
_OBJ_DIR = os.path.join(_HERE, '_build', 'obj')

SOURCES = ['common.cu', 'tmap.cu', 'nms.cu', 'roi_align.cu', 'layout.cu', 'warp.cu', 'tafa.cu', 'tafa_keyproj.cu', 'selsa.cu',
           'selsa_tc.cu', 'msra_gemm.cu', 'msra_overflow.cu', 'decode.cu', 'deform.cu']
# test-only kernels (include/vodagg_selftest.h): linked with the shared helpers into a SEPARATE library, so that
# libvodagg.so carries no test code
SELFTEST_LIB_PATH = os.path.join(_HERE, 'libvodagg_selftest.so')
SELFTEST_SOURCES = ['gemm_test.cu']
SELFTEST_SHARED = ['common.cu', 'tmap.cu']

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found; cannot build libvodagg.so')
    return nvcc


def _host_cxx():
    for c in ('/usr/bin/g++', shutil.which('g++')):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('g++ not found')


def _digest():
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ['../../include/vodagg.h', '../../include/vodagg_selftest.h']
    for n in names:
        p = os.path.join(CSRC, n)
        if os.path.isfile(p):
            h.update(n.encode())
            h.update(open(p, 'rb').read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force=False, verbose=False, probes=False):
    """Compile every .cu for sm_100a and link libvodagg.so (+ libvodagg_selftest.so, the test-only kernels). Returns the library path.

    ``probes=True`` (``python -m ...build --probes``): an experiment build with -DVOD_PROBES (environment-selected kernel variants
    and probe hooks) written to libvodagg_probes.so; the production library never reads the environment."""
    if probes:
        return _build_probes(verbose)
    stamp = os.path.join(_HERE, '_build', 'stamp')
    digest = _digest()
    if (not force and os.path.exists(LIB_PATH) and os.path.exists(SELFTEST_LIB_PATH) and os.path.exists(stamp)
            and open(stamp).read().strip() == digest):
        return LIB_PATH
    nvcc, cxx = _nvcc(), _host_cxx()
    os.makedirs(_OBJ_DIR, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(_OBJ_DIR, src.replace('.cu', '.o'))
        cmd = [nvcc, '-ccbin', cxx] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        if verbose and r.stderr:
            print(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = dict(zip(SOURCES + SELFTEST_SOURCES, ex.map(compile_one, SOURCES + SELFTEST_SOURCES)))
    for out, names in ((LIB_PATH, SOURCES), (SELFTEST_LIB_PATH, SELFTEST_SHARED + SELFTEST_SOURCES)):
        cmd = [nvcc, '-ccbin', cxx, '-shared', '-o', out] + [objs[n] for n in names] + ['-gencode', 'arch=compute_100a,code=sm_100a']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    with open(stamp, 'w') as f:
        f.write(digest)
    return LIB_PATH


def _build_probes(verbose=False):
    nvcc, cxx = _nvcc(), _host_cxx()
    # VOD_EXTRA_DEFINES="-DVOD_WARP_PIX=512 -DVOD_WARP_CH=16" VOD_PROBES_TAG=b python -m ...build --probes -> libvodagg_probes_b.so
    extra = os.environ.get('VOD_EXTRA_DEFINES', '').split()
    tag = os.environ.get('VOD_PROBES_TAG', '')
    obj_dir = os.path.join(_HERE, '_build', 'obj_probes' + tag)
    os.makedirs(obj_dir, exist_ok=True)
    out = os.path.join(_HERE, 'libvodagg_probes%s.so' % ('_' + tag if tag else ''))

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace('.cu', '.o'))
        r = subprocess.run([nvcc, '-ccbin', cxx] + NVCC_FLAGS + ['-DVOD_PROBES'] + extra + ['-c', os.path.join(CSRC, src), '-o', obj],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        return obj
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, '-ccbin', cxx, '-shared', '-o', out] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return out


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv, probes='--probes' in sys.argv))
