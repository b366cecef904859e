#!/usr/bin/env python
"""Group kernel vs the one-CTA-per-walker kernel (and the oracle on a sample): values and kernel time.

    python scripts/group_check.py [small|c4|c1] [nw]

Prints one JSON line per problem: max relative difference between the two device paths, against the oracle
on `--oracle` vectors, and the event-timed kernel duration of both paths.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else 'small'
    nw = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    n_oracle = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    if which == 'small':
        prob = synthetic.make_problem_small()
    elif which == 'c4':
        prob = synthetic.make_problem_c4()
    elif which == 'c1':
        prob = synthetic.make_problem()
    else:
        raise SystemExit(which)
    theta = synthetic.sample_prior_box(prob, nw, seed=5)
    if which == 'c4' and len(sys.argv) > 4 and sys.argv[4] == 'bench':
        sys.path.insert(0, ROOT)
        import bench
        theta = bench.initial_ensemble(prob, nw)
    lp = LogPosterior(prob)
    eng = lp.engine
    out = {'problem': which, 'nw': nw, 'group_supported': eng.group_supported}
    res = {}
    for name, mode in (('stream', False), ('group', True)):
        eng.set_group(mode)
        got = lp(theta)                       # warm-up + values
        eng.set_timing(True); eng.kernel_time()
        t0 = time.perf_counter()
        for _ in range(5):
            lp(theta)
        wall = (time.perf_counter() - t0) / 5
        ms, nl, nwk = eng.kernel_time()
        eng.set_timing(False)
        res[name] = got
        out[name + '_kernel_ms'] = ms / max(nl, 1)
        out[name + '_wall_ms'] = wall * 1e3
        out[name + '_evals_per_s_kernel'] = nw / (ms / max(nl, 1) / 1e3) if ms > 0 else None
    a, b = res['stream'], res['group']
    out['inf_sets_equal'] = bool(np.array_equal(np.isfinite(a), np.isfinite(b)))
    f = np.isfinite(a) & np.isfinite(b)
    out['n_finite'] = int(f.sum())
    out['max_rel_group_vs_stream'] = float(np.max(np.abs(a[f] - b[f]) / np.abs(a[f]))) if f.any() else None
    if n_oracle:
        from oracle_adapter import to_oracle
        want = to_oracle(prob).lnprob_batch(theta[:n_oracle])
        g = res['group'][:n_oracle]
        fo = np.isfinite(want)
        out['oracle_inf_sets_equal'] = bool(np.array_equal(np.isfinite(g), fo))
        out['max_rel_group_vs_oracle'] = float(np.max(np.abs(g[fo] - want[fo]) / np.abs(want[fo])))
    print(json.dumps(out), flush=True)
    lp.close()


if __name__ == '__main__':
    main()
