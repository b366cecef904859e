#!/usr/bin/env python
"""Kernel tuning helper: build libmcmcspec.so with different -D switches (here, on the CPU box), then on
the GPU box time every variant on the BASELINE shapes.

    python scripts/tune_variants.py build   name:DEF1=1,DEF2=0 ...     # -> build/variants/lib_<name>.so
    python scripts/tune_variants.py run [c1 c4 c5]                      # one JSON line per variant and shape
"""
import concurrent.futures as cf
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, 'mcmc_spec_b200', 'variants')


def build_one(spec):
    from mcmc_spec_b200 import build
    name, _, defs = spec.partition(':')
    out = os.path.join(VDIR, f'lib_{name}.so')
    build.build(force=True, defines=[d for d in defs.split(',') if d], out=out)
    return out


if __name__ == '__main__':
    if sys.argv[1] == 'build':
        os.makedirs(VDIR, exist_ok=True)
        with cf.ThreadPoolExecutor(4) as ex:
            for out in ex.map(build_one, sys.argv[2:]):
                print(out)
    else:
        which = sys.argv[2:] or ['c4']
        for lib in sorted(glob.glob(os.path.join(VDIR, 'lib_*.so'))):
            env = dict(os.environ, MCMCSPEC_LIB=lib)
            r = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'bench_configs.py')] + which,
                               env=env, capture_output=True, text=True)
            for line in r.stdout.splitlines():
                if line.startswith('{'):
                    d = json.loads(line)
                    print(json.dumps({'variant': os.path.basename(lib), 'config': d['config'][:12],
                                      'ms': round(d['kernel_ms_per_launch'], 4), 'evals_per_s': round(d['evals_per_s'])}), flush=True)
            if r.returncode:
                print(json.dumps({'variant': os.path.basename(lib), 'error': r.stderr[-400:]}), flush=True)
