"""In-tree build of libpasta_b200.so (the C-ABI library of include/pasta_b200.h) for sm_100a.

    python pasta-gan_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so sits in pasta-gan_b200/lib/ (git-ignored, but it
travels to the GPU box with the gpurun snapshot).  Replaces the reference's JIT plugin loader
(torch_utils/custom_ops.py:46-124): there is no md5-keyed cache and no fallback — if the library is
missing or stale the loader (_capi.py) rebuilds it here or fails loudly.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libpasta_b200.so')
STAMP = os.path.join(LIBDIR, 'libpasta_b200.stamp')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-ffp-contract=off',      # host code (pg_patch_geometry.cu) must round like OpenCV / numpy: no FMA contraction
              '--expt-relaxed-constexpr', '-I', INCLUDE]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def source_digest():
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h')))
    files.append(os.path.join(INCLUDE, 'pasta_b200.h'))
    for path in files:
        h.update(os.path.basename(path).encode())
        with open(path, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS[:-1]).encode())
    return h.hexdigest()


def is_current():
    try:
        with open(STAMP) as fh:
            return os.path.exists(LIB) and fh.read().strip() == source_digest()
    except OSError:
        return False


def nvcc_path():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def build(force=False, verbose=False, debug=False):
    """Compile every .cu under csrc/ for sm_100a and link libpasta_b200.so.  Returns the library path.
    ``debug``: a second library, lib/libpasta_b200_dbg.so, compiled with -DPG_DEBUG (per-CTA phase timestamps, load / store suppression and the
    loader experiments of tools/conv_timeline.py; results may be wrong by design) — never loaded unless PASTA_B200_LIB points at it."""
    if debug:
        return _build_debug(verbose)
    if not force and is_current():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = nvcc_path()
    objs = []

    def compile_one(src):
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    r = subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB] + objs + ['-lcudart_static', '-lpthread', '-ldl', '-lrt'],
                       capture_output=True, text=True)      # (the arch flag also on the link line: without it nvcc adds an empty default-arch fatbin entry)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    with open(STAMP, 'w') as fh:
        fh.write(source_digest())
    return LIB


def _build_debug(verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = nvcc_path()
    out = os.path.join(LIBDIR, 'libpasta_b200_dbg.so')
    cmd = [nvcc] + NVCC_FLAGS + ['-DPG_DEBUG', '-shared', '-o', out] + sources() + ['-lcudart_static', '-lpthread', '-ldl', '-lrt']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'debug build failed:\n{r.stdout}\n{r.stderr}')
    return out


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv, debug='--debug' in sys.argv))
