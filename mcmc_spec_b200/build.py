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


def build(force=False, verbose=False, defines=(), out=None):
    """Compile csrc/mcmcspec.cu -> libmcmcspec.so for sm_100a; returns the library path.
    ``defines`` / ``out``: tuning builds with extra -D macros into another file (scripts/tune_variants.py)."""
    lib = out or LIB
    deps = [SRC, os.path.join(ROOT, 'include', 'mcmcspec.h')]
    if (not force and os.path.exists(lib)
            and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in deps)):
        return lib
    cmd = [nvcc_path(), '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
           '-shared', '-Xcompiler', '-fPIC', '-cudart', 'static', '-I', os.path.join(ROOT, 'include'),
           '-o', lib, SRC] + ['-D' + d for d in defines]
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed building libmcmcspec.so')
    return lib


if __name__ == '__main__':
    print(build(force=True, verbose='-v' in sys.argv))
