"""GPU parity of every exported prior flag combination (ms_config.dist_fit / rad_prior / ext_flag).

logprior has four 6-parameter and four 8-parameter flavours (mft6.py:1212-1327, :1329-1454) selected by
``dist_fit`` and ``ext`` (= logposterior's ``a``), each with or without the isochrone radius prior.  ``main``
only ever calls a = True (mft6.py:3689), but the flags are part of the boundary, so each is held to the oracle:
binaries and triples x dist_fit x rad_prior x a, 2048 parameter vectors each (edge cases included), bar 1e-9
relative with identical -inf sets.  With ext == False the 6-parameter dist_fit branch bounds and weighs
p0[nspec:2*nspec] -- the Av slot and R1 -- as its radii (mft6.py:1219-1221); the sample spreads the Av column
over that box so both outcomes of every test occur.
"""
import dataclasses

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 2048


def _thetas(prob, seed):
    from mcmc_spec_b200 import synthetic
    if prob.nspec == 2:
        th = synthetic.sample_prior_box(prob, N, seed=seed)
    else:
        th = synthetic.sample_prior_box_triple(prob, N, seed=seed)
    rng = np.random.default_rng(seed + 1)
    k = np.arange(N) % 4 == 1
    th[k, prob.nspec] = rng.uniform(0.0, 1.7, k.sum())          # Av slot over the R1 box (ext == False)
    # a cloud near the truth, where the radius prior does not dwarf the likelihood
    m = np.arange(N) % 4 == 2
    th[m] = np.asarray(prob.truth) * (1 + 0.01 * rng.standard_normal((m.sum(), th.shape[1])))
    th[m, prob.nspec] = np.abs(th[m, prob.nspec])
    return th


@pytest.mark.parametrize('nspec', [2, 3])
@pytest.mark.parametrize('dist_fit', [True, False])
@pytest.mark.parametrize('rad_prior', [True, False])
@pytest.mark.parametrize('a', [True, False])
def test_prior_flag_combination(native_lib, nspec, dist_fit, rad_prior, a):
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    make = synthetic.make_problem_small if nspec == 2 else synthetic.make_problem_triple_small
    prob = dataclasses.replace(make(dist_fit=dist_fit, rad_prior=rad_prior), av=a)
    if nspec == 3 and dist_fit and not rad_prior:
        # mft6.py:1381-1393: this branch only returns inside `if rad_prior == True` -> None -> the reference
        # raises in logposterior; the oracle raises ValueError and the library refuses with MS_ERR_UNSUPPORTED
        from mcmc_spec_b200 import native
        with pytest.raises(ValueError):
            to_oracle(prob).lnprob(np.asarray(prob.truth))
        lp = LogPosterior(prob)
        with pytest.raises(native.NativeError):
            lp(np.asarray(prob.truth)[None, :])
        lp.close()
        return
    theta = _thetas(prob, seed=40 + 8 * nspec + 4 * dist_fit + 2 * rad_prior + a)
    want = to_oracle(prob).lnprob_batch(theta)
    lp = LogPosterior(prob)
    got = lp(theta)
    lp.close()
    assert not np.any(np.isnan(got))
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    f = np.isfinite(want)
    assert 64 <= f.sum() < N                                     # both finite and -inf vectors are present
    rel = np.abs(got[f] - want[f]) / np.abs(want[f])
    assert rel.max() <= 1e-9, f"max rel err {rel.max():.3e}"


def test_device_bayestar_rule_documented_cases(native_lib):
    """walker_prep_kernel's Av prior against the dustmaps rule written out by hand (tests/test_published_known_answers.py
    cites it): exactly on a distance slice, half-way in DM between two slices, nearer than the first slice (linear in
    distance), beyond the last slice -- with every other prior term switched off."""
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    prob = synthetic.make_problem_small(rad_prior=False)
    prob.prior = [0.0] * len(prob.prior)
    edges, samples = prob.dust_dm_edges, prob.dust_samples
    k10 = int(round((10.0 - edges[0]) / (edges[1] - edges[0])))
    cases = {
        1000.0: samples[:, k10],
        10 ** ((10.0 + 0.0625) / 5 + 1): 0.5 * (samples[:, k10] + samples[:, k10 + 1]),
        0.5 * 10 ** (edges[0] / 5 + 1): 0.5 * samples[:, 0],
        2900.0: None,
    }
    theta = np.tile(np.asarray(prob.truth, dtype=float), (len(cases), 1))
    for k, d in enumerate(cases):
        theta[k, 5] = 1.0 / d
    lp = LogPosterior(prob)
    got = lp.parts(theta)['lnprior']
    lp.close()
    for k, (d, e) in enumerate(cases.items()):
        if e is None:
            continue
        e = e * 3.1 * 0.884
        want = -0.5 * ((theta[k, 2] - e.mean()) / (e.std() if e.std() else 0.05)) ** 2
        assert got[k] == pytest.approx(want, rel=1e-11, abs=1e-13), d
