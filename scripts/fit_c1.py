#!/usr/bin/env python
"""Config C1 end to end at the reference's own sizes (param_koi2298.txt: 150 optimiser walks x 400 steps,
best third = 50 walkers, 301 burn-in + 15001 production steps, dumps every 100) on the synthetic grid:
recovery of a 3850 K + 3425 K binary (the truths of the reference's synth_spec recipe, mft6.py:3599-3605).

    python scripts/fit_c1.py [outdir] [--nsteps N] [--no-converge]

Prints one JSON line: wall time of each stage, evaluations/s, acceptance fraction, posterior medians and
16/84 percentiles against the truth.  This is the workload a user of the reference actually runs.
"""
import argparse
import json
import os
import sys
import tempfile
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np   # noqa: E402
import torch         # noqa: E402

from mcmc_spec_b200 import fit, synthetic                 # noqa: E402
from mcmc_spec_b200.posterior import LogPosterior         # noqa: E402

if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('outdir', nargs='?', default=None)
    ap.add_argument('--nwalk', type=int, default=150)
    ap.add_argument('--nstep', type=int, default=400)
    ap.add_argument('--nburn', type=int, default=301)
    ap.add_argument('--nsteps', type=int, default=15001)
    ap.add_argument('--no-converge', action='store_true', help='run all nsteps (skip the autocorrelation stop)')
    ap.add_argument('--fast', action='store_true', help='sample in the bracket-only mode (ms_set_fast)')
    args = ap.parse_args()
    out = args.outdir or tempfile.mkdtemp(prefix='fit_c1_')
    # truth on the synthetic isochrone (the reference's recipe takes its radii from the real MIST track)
    prob = synthetic.make_problem(truth=synthetic.consistent_truth(3850.0, 3425.0, 0.12, 2.0732e-3))
    lp = LogPosterior(prob)
    plx, plx_err = prob.prior[-2], prob.prior[-1]
    fp = SimpleNamespace(nwalk=args.nwalk, nstep=args.nstep, plx=plx, plx_err=plx_err, fname='synth_3850_3425',
                         nsteps=args.nsteps, nburn=args.nburn)
    n0 = lp.engine.launch_count
    t0 = time.perf_counter()
    fit.run_fit(lp, fp, out, optimize=True, emcee=False, seed=2298)
    torch.cuda.synchronize()
    t_opt = time.perf_counter() - t0
    n_opt_props = sum(sum(1 for _ in open(os.path.join(out, f'chisq{k}.txt'))) for k in range(args.nwalk))
    msgs = []
    lp.engine.set_fast(args.fast)
    t1 = time.perf_counter()
    if args.no_converge:
        from mcmc_spec_b200 import optimizer
        from mcmc_spec_b200.ensemble import CudaBackend, EnsembleSampler
        p0 = optimizer.select_walkers(np.genfromtxt(os.path.join(out, 'optimize_res.txt')),
                                      np.genfromtxt(os.path.join(out, 'optimize_cs.txt')))
        s = EnsembleSampler(len(p0), 6, CudaBackend(lp.engine), seed=2298)
        samples = fit.run_emcee(out, fp.fname, s, p0, fp.nsteps, fp.nburn, nthin=100, check_convergence=False, log=msgs.append)
    else:
        samples = fit.run_fit(lp, fp, out, optimize=False, emcee=True, seed=2298, nthin=100, log=msgs.append)
    torch.cuda.synchronize()
    t_mc = time.perf_counter() - t1
    nw = args.nwalk // 3
    nw -= nw % 2
    kept = samples.shape[0] // nw
    n_mc_evals = nw * (args.nburn + kept + 1)
    q = np.percentile(samples, [16, 50, 84], axis=0)
    print(json.dumps({
        "config": "C1: 3850 K + 3425 K synthetic binary, koi2298 settings (N=1945, 135000-px rows, 26x4 grid), "
                  "150 optimiser walks x 400 steps -> 50 walkers x (301 + 15001) steps",
        "sampler_mode": "bracket-only" if args.fast else "streaming", "optimizer_s": round(t_opt, 2), "optimizer_proposals": n_opt_props,
        "optimizer_evals_per_s": round(n_opt_props / t_opt),
        "sampler_s": round(t_mc, 2), "sampler_steps_kept": kept, "sampler_evals": n_mc_evals,
        "sampler_evals_per_s": round(n_mc_evals / t_mc), "sampler_ms_per_iteration": round(1e3 * t_mc / (args.nburn + kept), 4),
        "acceptance": msgs[0] if msgs else None, "launches": int(lp.engine.launch_count - n0),
        "truth": [float(v) for v in prob.truth],
        "p16": [float(v) for v in q[0]], "median": [float(v) for v in q[1]], "p84": [float(v) for v in q[2]],
        "outdir_files": len(os.listdir(out))}))
    lp.close()
