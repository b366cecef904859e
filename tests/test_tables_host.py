"""Host logic: the table builder + the device arithmetic plan reproduce the oracle (CPU only)."""
import numpy as np
import pytest

from mcmc_spec_b200 import synthetic
from mcmc_spec_b200.tables import TILE, build_tables, orthonormal_quadratic_basis, trapz_weights, ccm89_k
from oracle import thirdparty as tp
import tables_emulator


def test_trapz_weights_and_ccm89_k():
    x = np.sort(np.random.default_rng(0).uniform(0, 10, 57)); y = np.sin(x)
    assert trapz_weights(x) @ y == pytest.approx(np.trapezoid(y, x), rel=1e-14)
    lam = np.linspace(3100, 29000, 777)
    assert np.allclose(ccm89_k(lam), tp.ccm89(lam, 1.0), rtol=1e-15, atol=0)


def test_orthonormal_basis_equals_polynomial_fit():
    rng = np.random.default_rng(1)
    x = np.sort(rng.uniform(-1, 1, 400)); x[0], x[-1] = -1.0, 1.0
    y = 1 + 0.05 * rng.standard_normal(400) + 0.1 * x - 0.2 * x * x
    B = orthonormal_quadratic_basis(x)
    q = np.stack([np.full_like(x, B[0, 0]), B[1, 0] + B[1, 1] * x, B[2, 0] + B[2, 1] * x + B[2, 2] * x * x])
    assert np.allclose(q @ q.T, np.eye(3), atol=1e-14)
    fit = (q @ y) @ q
    ref = np.polynomial.Polynomial.fit(x, y, 2)(x)
    assert np.allclose(fit, ref, rtol=1e-13)


def test_table_shapes_and_region(small_problem):
    t = build_tables(small_problem)
    assert t.n_tiles == -(-t.n_reg // TILE)
    assert t.rebin_hi.min() >= 1 and t.rebin_hi.max() < t.n_reg and np.all(np.diff(t.rebin_hi) >= 0)
    assert np.all((t.rebin_t >= 0) & (t.rebin_t <= 1))
    # region covers data and every filter, like make_composite's get_spec cut (mft6.py:687)
    lo, hi = t.wl_reg[0], t.wl_reg[-1]
    assert lo <= small_problem.data[0].min() * 1e4 and hi >= small_problem.data[0].max() * 1e4
    assert lo <= small_problem.tmi + 1.0 and hi >= small_problem.tma - 1.0
    assert t.bytes_per_eval() == 2 * 4 * t.n_reg * 8 + 8 * 6 + 8
    for pk in (t.contrast, t.phot):
        assert pk.weights.size % TILE == 0 and np.all(pk.tile_hi <= t.n_tiles)


def test_device_plan_matches_oracle(small_problem, small_oracle):
    t = build_tables(small_problem)
    theta = synthetic.sample_prior_box(small_problem, 48, seed=21)
    want = small_oracle.lnprob_batch(theta)
    got = np.array([tables_emulator.lnprob(t, th) for th in theta])
    assert np.array_equal(np.isfinite(want), np.isfinite(got))
    f = np.isfinite(want)
    assert f.sum() >= 40
    assert np.max(np.abs(got[f] - want[f]) / np.abs(want[f])) < 1e-11


@pytest.mark.parametrize('dist_fit', [True, False])
@pytest.mark.parametrize('rad_prior', [True, False])
@pytest.mark.parametrize('av', [True, False])
def test_device_plan_prior_flags(dist_fit, rad_prior, av):
    """Every (dist_fit, rad_prior, a) branch of the 6-parameter logprior (mft6.py:1212-1327), including
    ext == False where the reference slices p0[nspec:2*nspec] as its radii (mft6.py:1219-1221)."""
    import dataclasses
    from oracle_adapter import to_oracle
    prob = dataclasses.replace(synthetic.make_problem_small(dist_fit=dist_fit, rad_prior=rad_prior), av=av)
    t = build_tables(prob)
    theta = flag_thetas(prob, 96, seed=33)
    want = to_oracle(prob).lnprob_batch(theta)
    got = np.array([tables_emulator.lnprob(t, th) for th in theta])
    assert np.array_equal(np.isfinite(want), np.isfinite(got))
    f = np.isfinite(want)
    assert 10 <= f.sum() < len(theta)
    assert np.max(np.abs(got[f] - want[f]) / np.abs(want[f])) < 1e-11


def flag_thetas(prob, n, seed):
    """Prior-box sample whose Av column also covers [0.05, 1.5] x the radius bounds: with ext == False the Av slot
    is what the reference bounds and weighs as `R1` (mft6.py:1219-1227), so both signs of each test are hit."""
    th = synthetic.sample_prior_box(prob, n, seed=seed)
    rng = np.random.default_rng(seed + 1)
    k = np.arange(n) % 4 == 1
    th[k, 2] = rng.uniform(0.0, 1.7, k.sum())
    return th


def test_unsorted_filter_curve_is_stably_sorted():
    # lp600-like digitised curve: descending steps and a duplicated abscissa
    w = np.array([6000., 6100., 6050., 6200., 6200., 6300.]); tr = np.array([0., .5, .4, .8, .7, 0.])
    p = synthetic.make_problem_small(contrast_curves=[(w, tr), (w + 500, tr)], contrast_names=['a', 'b'])
    t = build_tables(p)
    from oracle_adapter import to_oracle
    o = to_oracle(p)
    th = p.truth.copy()
    assert tables_emulator.lnprob(t, th) == pytest.approx(o.lnprob(th), rel=1e-11)
