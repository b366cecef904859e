"""In-tree build of libmcmcspec.so (nvcc, sm_100a).  `python -m mcmc_spec_b200.build`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, 'csrc', 'mcmcspec.cu')
LIB = os.path.join(HERE, 'libmcmcspec.so')


def nvcc_path():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def build(force=False, verbose=False, defines=(), out=None, split=True):
    """Compile csrc/mcmcspec.cu -> libmcmcspec.so for sm_100a; returns the library path.

    ``split``: the source is compiled four times in parallel (-DMS_TU=1/2/3/4: the streaming kernel's ~60
    instantiations are spread over three translation units, the group kernel is the fourth) and the objects are linked; ``split=False`` is the
    single-unit build (-DMS_TU=0).  ``defines`` / ``out``: tuning builds with extra -D macros into another file
    (scripts/tune_variants.py)."""
    lib = out or LIB
    deps = [SRC, os.path.join(HERE, 'csrc', 'group_kernel.cuh'), os.path.join(ROOT, 'include', 'mcmcspec.h')]
    if (not force and os.path.exists(lib)
            and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in deps)):
        return lib
    common = [nvcc_path(), '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include')] + ['-D' + d for d in defines]
    if verbose:
        common[1:1] = ['-Xptxas', '-v']

    def run(cmd):
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError('nvcc failed building libmcmcspec.so')

    if not split:
        run(common + ['-shared', '-cudart', 'static', '-DMS_TU=0', '-o', lib, SRC])
        return lib
    objdir = os.path.join(HERE, 'build', os.path.basename(lib) + '.obj')
    os.makedirs(objdir, exist_ok=True)
    objs = [os.path.join(objdir, f'tu{k}.o') for k in (1, 2, 3, 4)]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(4) as ex:
        list(ex.map(lambda ko: run(common + ['-c', '-Xfatbin', '-compress-all', f'-DMS_TU={ko[0]}', '-o', ko[1], SRC]),
                    zip((1, 2, 3, 4), objs)))
    run([nvcc_path(), '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-cudart', 'static', '-Xcompiler', '-fPIC',
         '-o', lib] + objs)
    shutil.rmtree(objdir, ignore_errors=True)          # the objects are only needed for the link
    try:
        os.rmdir(os.path.dirname(objdir))
    except OSError:
        pass
    return lib


if __name__ == '__main__':
    print(build(force=True, verbose='-v' in sys.argv))
