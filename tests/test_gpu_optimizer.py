"""GPU: the optimiser stage (fit_spec's chi-squares) and the mft6_nospec flavour (f2)
against the oracle."""
import dataclasses

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _starts(prob, n, seed):
    from mcmc_spec_b200 import optimizer as opt
    rng = np.random.default_rng(seed)
    th = opt.draw_starts(n, float(min(prob.temps)), float(max(prob.temps)), prob.prior[-2], prob.prior[-1], rng)
    return th


def test_fit_spec_chisq_matches_oracle(small_problem, small_oracle, native_lib):
    from mcmc_spec_b200.optimizer import GpuChisq
    from mcmc_spec_b200.posterior import LogPosterior
    from opt_reference import OracleChisq
    lp = LogPosterior(small_problem)
    gpu, orc = GpuChisq(lp), OracleChisq(small_problem, small_oracle)
    theta0 = _starts(small_problem, 24, 1)
    g_cs, g_sets = gpu.start(theta0)
    o_cs, o_sets = orc.start(theta0)
    assert np.allclose(g_sets, o_sets, rtol=1e-12)                 # the one-off renormalised spectra
    assert np.allclose(g_cs, o_cs, rtol=1e-9)                      # init_cs, mft6.py:891-930
    rng = np.random.default_rng(2)
    theta = theta0[rng.integers(0, 24, 200)] * (1 + 0.03 * rng.standard_normal((200, 6)))
    theta[:, 2] = np.abs(theta[:, 2]); theta[:5, 2] = 0.0          # Av = 0: extinction skipped (mft6.py:1001)
    theta[:, 0:2] = np.clip(theta[:, 0:2], min(small_problem.temps) + 1, max(small_problem.temps) - 1)
    idx = rng.integers(0, 24, 200).astype(np.int32)
    rs = 0.1 * theta0[idx, 3:5]
    got = gpu.test(theta, idx, rs)
    want = orc.test(theta, idx, rs)
    assert np.all(np.isfinite(want)) and np.max(np.abs(got - want) / np.abs(want)) < 1e-9
    # undefined forward model -> +inf (never wins `test_cs < chi`)
    bad = theta[:2].copy(); bad[0, 0] = 9000.0; bad[1, 5] = -1.0
    assert np.all(gpu.test(bad, idx[:2], rs[:2]) == np.inf)
    lp.close()


def test_optimizer_walk_matches_oracle_walk(small_problem, small_oracle, native_lib, tmp_path):
    """Same random draws fed to the GPU-backed and the oracle-backed walk: same decisions, same optimum."""
    from mcmc_spec_b200 import optimizer as opt
    from mcmc_spec_b200.posterior import LogPosterior
    from opt_reference import OracleChisq
    lp = LogPosterior(small_problem)
    plx, plx_err = small_problem.prior[-2], small_problem.prior[-1]
    g_par, g_cs = opt.optimize_fit(lp, 8, 10, plx, plx_err, seed=4, dirname=str(tmp_path / 'run'))
    o_par, o_cs = opt.optimize_fit(lp, 8, 10, plx, plx_err, seed=4, backend=OracleChisq(small_problem, small_oracle))
    assert np.allclose(g_par, o_par, rtol=1e-12) and np.allclose(g_cs, o_cs, rtol=1e-9)
    res = np.genfromtxt(tmp_path / 'run' / 'optimize_res.txt')      # the reference's file formats
    cs = np.genfromtxt(tmp_path / 'run' / 'optimize_cs.txt')
    assert res.shape == (8, 6) and np.allclose(res, g_par) and np.allclose(cs, g_cs)
    assert (tmp_path / 'run' / 'params0.txt').exists() and (tmp_path / 'run' / 'chisq7.txt').exists()
    sel = opt.select_walkers(res, cs)
    assert sel.shape == (2, 6)
    lp.close()


def test_nospec_flavour_matches_oracle(small_problem, native_lib):
    """mft6_nospec.py: contrast + photometry only, parallax box up to 1/100."""
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    prob = dataclasses.replace(small_problem, nospec=True)
    lp = LogPosterior(prob)
    theta = synthetic.sample_prior_box(prob, 256, seed=12)
    theta[20, 5] = 1 / 50.0                                         # inside mft6.py's box, outside mft6_nospec.py's
    got, want = lp(theta), to_oracle(prob).lnprob_batch(theta)
    assert got[20] == -np.inf and want[20] == -np.inf
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    f = np.isfinite(want)
    assert np.max(np.abs(got[f] - want[f]) / np.abs(want[f])) < 1e-9
    full = LogPosterior(small_problem)
    assert not np.allclose(full(theta)[f], got[f])                   # the spectral term really is gone
    full.close(); lp.close()


def test_nospec_optimizer_chisq_matches_oracle(small_problem, native_lib):
    """fit_spec of mft6_nospec.py: the optimiser's chi-square without the spectral term (contrast + photometry +
    the chi-square priors)."""
    from mcmc_spec_b200.optimizer import GpuChisq
    from mcmc_spec_b200.posterior import LogPosterior
    from opt_reference import OracleChisq
    from oracle_adapter import to_oracle
    prob = dataclasses.replace(small_problem, nospec=True)
    lp = LogPosterior(prob)
    gpu, orc = GpuChisq(lp), OracleChisq(prob, to_oracle(prob))
    theta0 = _starts(prob, 16, 5)
    g_cs, _ = gpu.start(theta0)
    o_cs, _ = orc.start(theta0)
    assert np.allclose(g_cs, o_cs, rtol=1e-9)
    rng = np.random.default_rng(6)
    theta = theta0[rng.integers(0, 16, 64)] * (1 + 0.03 * rng.standard_normal((64, 6)))
    theta[:, 2] = np.abs(theta[:, 2])
    theta[:, 0:2] = np.clip(theta[:, 0:2], min(prob.temps) + 1, max(prob.temps) - 1)
    idx = rng.integers(0, 16, 64).astype(np.int32)
    rs = 0.1 * theta0[idx, 3:5]
    got, want = gpu.test(theta, idx, rs), orc.test(theta, idx, rs)
    assert np.all(np.isfinite(want)) and np.max(np.abs(got - want) / np.abs(want)) < 1e-9
    full = GpuChisq(LogPosterior(small_problem))
    full.start(theta0)
    assert not np.allclose(full.test(theta, idx, rs), got)          # the spectral term really is gone
    lp.close()


def test_triple_star_optimizer(native_lib, tmp_path):
    """fit_spec / optimize_fit for nspec = 3 (mft6.py:920-922, :952-990, :1044-1050, :1718-1741): 8-parameter walks,
    three radius priors, the tertiary ratio forced below the secondary's; chi-squares against the oracle and a whole
    walk with the same random draws against the oracle-backed walk."""
    from mcmc_spec_b200 import optimizer as opt, synthetic
    from mcmc_spec_b200.optimizer import GpuChisq
    from mcmc_spec_b200.posterior import LogPosterior
    from opt_reference import OracleChisq
    from oracle_adapter import to_oracle
    prob = synthetic.make_problem_triple_small()
    orc_p = to_oracle(prob)
    lp = LogPosterior(prob)
    plx, plx_err = prob.prior[-2], prob.prior[-1]
    rng = np.random.default_rng(3)
    theta0 = opt.draw_starts(12, float(min(prob.temps)), float(max(prob.temps)), plx, plx_err, rng, nspec=3)
    assert theta0.shape == (12, 8) and np.all(theta0[:, 2] <= theta0[:, 1]) and np.all(theta0[:, 1] <= theta0[:, 0])
    gpu, orc = GpuChisq(lp), OracleChisq(prob, orc_p)
    g_cs, g_sets = gpu.start(theta0)
    o_cs, o_sets = orc.start(theta0)
    assert np.allclose(g_sets, o_sets, rtol=1e-12) and np.allclose(g_cs, o_cs, rtol=1e-9)
    theta = theta0[rng.integers(0, 12, 48)] * (1 + 0.02 * rng.standard_normal((48, 8)))
    theta[:, 3] = np.abs(theta[:, 3])
    theta[:, 0:3] = np.clip(theta[:, 0:3], min(prob.temps) + 1, max(prob.temps) - 1)
    idx = rng.integers(0, 12, 48).astype(np.int32)
    rs = 0.1 * theta0[idx, 4:7]
    got, want = gpu.test(theta, idx, rs), orc.test(theta, idx, rs)
    assert np.all(np.isfinite(want)) and np.max(np.abs(got - want) / np.abs(want)) < 1e-9
    g_par, g_chi = opt.optimize_fit(lp, 6, 6, plx, plx_err, seed=8, dirname=str(tmp_path / 'run'))
    o_par, o_chi = opt.optimize_fit(lp, 6, 6, plx, plx_err, seed=8, backend=OracleChisq(prob, orc_p))
    assert g_par.shape == (6, 8) and np.allclose(g_par, o_par, rtol=1e-12) and np.allclose(g_chi, o_chi, rtol=1e-9)
    assert np.genfromtxt(tmp_path / 'run' / 'optimize_res.txt').shape == (6, 8)
    lp.close()
