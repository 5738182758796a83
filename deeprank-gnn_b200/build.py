"""Build libdrgnn.so (the C-ABI library, ``include/drgnn.h``) in-tree with nvcc for sm_100a.

``python -m deeprank_gnn_b200.build`` or ``build()``; objects go to ``csrc/build/`` and the
shared library to the package directory, both git-ignored (the built ``.so`` still travels
to the GPU box with the working-tree snapshot).
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, 'build')
LIB = os.path.join(HERE, 'libdrgnn.so')
ROOT = os.path.dirname(HERE)

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include')]


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: libdrgnn.so cannot be built (set NVCC=/path/to/nvcc)')


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    hdrs.append(os.path.join(ROOT, 'include', 'drgnn.h'))
    return hdrs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


LAST = None      # record of the last build() of this process (see build_record)


def build_record():
    """What the last ``build()`` of this process did and which library is in place: sources compiled here (empty =
    the objects that travelled with the working tree were newer than every source and header), whether the
    library was re-linked, nvcc's release line, size / sha256 prefix / mtime of ``libdrgnn.so``.  bench.py prints it
    as ``native`` so a reader can tell a box that recompiled from one that reused the shipped binary."""
    import hashlib
    rec = dict(LAST or {'compiled': None, 'linked': None, 'nvcc': None})
    if os.path.exists(LIB):
        with open(LIB, 'rb') as f:
            blob = f.read()
        rec.update(so=os.path.relpath(LIB, ROOT), so_bytes=len(blob), so_sha256_16=hashlib.sha256(blob).hexdigest()[:16],
                   so_mtime=int(os.path.getmtime(LIB)))
    rec['flags'] = ' '.join(NVCC_FLAGS[:5])
    return rec


def build(force=False, verbose=False):
    """Compile every ``csrc/*.cu`` for sm_100a and link ``libdrgnn.so``.  Returns the path."""
    global LAST
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _deps()
    jobs = []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        if force or _stale(obj, [src] + hdrs):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get('DRGNN_NVCC_EXTRA', '').split() + ['-c', src, '-o', obj]
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(compile_one, jobs))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + '.o') for s in sources()]
    try:
        ver = [ln for ln in subprocess.run([nvcc, '--version'], capture_output=True, text=True).stdout.splitlines()
               if 'release' in ln]
    except OSError:
        ver = []
    LAST = {'compiled': [os.path.basename(j[0]) for j in jobs], 'linked': False, 'nvcc': ver[0].strip() if ver else None}
    if force or jobs or _stale(LIB, objs):
        LAST['linked'] = True
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-cudart', 'static']
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link of libdrgnn.so failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
