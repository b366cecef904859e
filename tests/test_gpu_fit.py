"""GPU: main's -o/-e flow end to end (optimiser -> best third -> burn-in -> production -> reference file formats)."""
from types import SimpleNamespace

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_run_fit_end_to_end(small_problem, small_oracle, native_lib, tmp_path):
    from mcmc_spec_b200 import fit
    from mcmc_spec_b200.posterior import LogPosterior
    lp = LogPosterior(small_problem)
    plx, plx_err = small_problem.prior[-2], small_problem.prior[-1]
    fp = SimpleNamespace(nwalk=48, nstep=6, plx=plx, plx_err=plx_err, fname='koi', nsteps=40, nburn=12)
    run = tmp_path / 'run'
    msgs = []
    samples = fit.run_fit(lp, fp, str(run), optimize=True, emcee=True, seed=3, nthin=10, log=msgs.append)
    assert samples.shape == (40 * 16, 6)                                  # best third of 48 walks, 40 kept steps
    assert np.allclose(np.loadtxt(run / 'samples.txt'), samples)
    assert np.genfromtxt(run / 'optimize_res.txt').shape == (48, 6)
    assert (run / 'koi_0_burnin.txt').exists() and (run / 'koi_10_burnin.txt').exists()
    assert (run / 'koi_30_results.txt').exists() and (run / 'koi_autocorr.txt').exists()
    # the chain never leaves the support of the posterior, and its log-probs are the oracle's
    chain = samples.reshape(16, 40, 6)
    last = chain[:, -1]
    got = lp(last)
    want = small_oracle.lnprob_batch(last)
    assert np.all(np.isfinite(got)) and np.max(np.abs(got - want) / np.abs(want)) <= 1e-9
    assert msgs and 'acceptance' in msgs[0]
    # -e only: re-reads the optimiser files, as the reference does (mft6.py:3668)
    again = fit.run_fit(lp, fp, str(run), optimize=False, emcee=True, seed=3, nthin=10, log=msgs.append)
    assert again.shape == samples.shape
    lp.close()
