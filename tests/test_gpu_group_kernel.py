"""The group kernel (lnprob_group_kernel: cell-mates share one TMA-staged stream of their grid rows, lane = walker)
against the oracle and against the one-CTA-per-walker kernel, on the cases where its own code paths differ: exact
median over odd / even / tiny / ragged sizes and many equal keys (per-lane histogram + key lists), partial groups,
walkers outside the prior box, both kernel variants (4 / 8 walker warps), intermediate outputs, determinism.

Bar: identical -inf sets, 1e-9 relative against the oracle (tolerance of north_star), 1e-11 between the two device
paths (they differ in the summation order of the band integrals and the moments).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    f = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), f)
    assert not np.any(np.isnan(a))
    return float(np.max(np.abs(a[f] - b[f]) / np.abs(b[f]))) if f.any() else 0.0


def _both(lp, theta):
    lp.engine.set_group(True)
    g = lp(theta)
    g2 = lp(theta)
    lp.engine.set_group(False)
    s = lp(theta)
    assert np.array_equal(g, g2)                                 # deterministic (grouping is not: atomics order the cells)
    return g, s


@pytest.mark.parametrize('n_data', [1, 2, 7, 63, 64, 255, 256, 300, 301, 1025])
def test_median_sizes_through_the_group_kernel(native_lib, n_data):
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    p = synthetic.make_problem_small(n_data=n_data, seed=40 + n_data)
    lp = LogPosterior(p)
    assert lp.engine.group_supported
    theta = synthetic.sample_prior_box(p, 200, seed=n_data)
    theta[100:] = np.asarray(p.truth) * (1 + 0.003 * np.random.default_rng(n_data).standard_normal((100, 6)))   # one dense cell
    theta[100:, 2] = np.abs(theta[100:, 2])
    g, s = _both(lp, theta)
    lp.close()
    assert _rel(g, to_oracle(p).lnprob_batch(theta)) <= 1e-9
    assert _rel(g, s) <= 1e-11


def test_equal_keys_and_huge_av_through_the_group_kernel(native_lib):
    """Flat rows: every model value equal (Av = 0) or a smooth ramp (Av > 0); Av far beyond the polynomial's range
    takes the exp2 path of the whole warp."""
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    p = synthetic.make_problem_small(n_data=200, seed=77)
    p.grid[...] = 1.0e6
    lp = LogPosterior(p)
    theta = synthetic.sample_prior_box(p, 96, seed=5)
    theta[:16, 2] = 0.0
    theta[16:24, 2] = 9.5                                         # |Av x| >> 0.02
    g, s = _both(lp, theta)
    lp.close()
    assert _rel(g, to_oracle(p).lnprob_batch(theta)) <= 1e-9
    assert _rel(g, s) <= 1e-11


def test_parts_and_model_through_the_group_kernel(small_problem, small_oracle, native_lib):
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    lp = LogPosterior(small_problem)
    lp.engine.set_group(True)
    theta = synthetic.sample_prior_box(small_problem, 80, seed=9)
    parts = lp.parts(theta)
    model = lp.model_spectrum(theta[16:24])
    for k in range(80):
        ref = {}
        v = small_oracle.lnprob(theta[k], ref)
        if not np.isfinite(v):
            assert parts['lnprob'][k] == -np.inf and parts['status'][k] != 0
            continue
        assert parts['lnprior'][k] == pytest.approx(ref['lp'], rel=1e-12)
        assert parts['chi2_spec'][k] == pytest.approx(ref['chi2_spec'], rel=1e-11)
        assert np.allclose(parts['contrast'][k], ref['contrast'], rtol=1e-12)
        assert np.allclose(parts['phot'][k], ref['phot'], rtol=1e-13)
        assert np.allclose(parts['logg'][k], ref['logg'], rtol=1e-14)
        if 16 <= k < 24:
            assert np.allclose(model[k - 16], ref['model'], rtol=1e-13)
    lp.close()


def test_both_kernel_variants_give_the_same_bits(native_lib):
    """Few groups -> eight walker warps per group on one CTA per SM, many groups -> four on two CTAs per SM; a
    walker's value must not depend on the variant, the batch it arrives in or its companions in the group (the
    sharded chain then walks the same path on any number of GPUs)."""
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    p = synthetic.make_problem_small()
    lp = LogPosterior(p)
    lp.engine.set_group(True)
    rng = np.random.default_rng(3)
    theta = np.asarray(p.truth) * (1 + 0.01 * rng.standard_normal((16384, 6)))
    theta[:, 2] = np.abs(theta[:, 2])
    big = lp(theta)
    big = lp(theta)                                               # second call: the variant follows the group count seen
    small = np.concatenate([lp(theta[k:k + 1024]) for k in range(0, 16384, 1024)])
    small = np.concatenate([lp(theta[k:k + 1024]) for k in range(0, 16384, 1024)])
    perm = rng.permutation(16384)
    shuffled = lp(theta[perm])
    lp.close()
    assert np.array_equal(big, small)
    assert np.array_equal(big[perm], shuffled)
