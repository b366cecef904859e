"""Posterior-predictive sweep (mft6.py:785-816, :2486-2593) on the GPU vs the oracle's plot=True branch."""
import numpy as np
import pytest

from mcmc_spec_b200 import synthetic
from mcmc_spec_b200.passbands import PhotBand

pytestmark = pytest.mark.gpu


def _bands(prob):
    """A Kepler-like energy curve partly outside the consumed region's blue end is avoided: both bands
    sit inside the region; the Gaia-like band is a photon-type library filter with a Vega zero flux."""
    wl = np.asarray(prob.wl)
    lo, hi = wl[0], wl[-1]
    mid = 0.5 * (lo + hi)
    kep = synthetic.synthetic_passband(mid - 0.05 * (hi - lo), 0.30 * (hi - lo), n=317, ripple_seed=41)
    gw, gt = synthetic.synthetic_passband(mid + 0.02 * (hi - lo), 0.45 * (hi - lo), n=253, ripple_seed=42)
    gaia = PhotBand('Gaia_G', gw, gt, 'photon', 'vega', vega_zero_flux=2.5e-9)
    return kep, gaia


def test_component_sweep_matches_oracle_plot_branch(small_problem, small_oracle, native_lib, tmp_path):
    from mcmc_spec_b200.posterior import LogPosterior
    from mcmc_spec_b200 import predictive
    from oracle import mft6_oracle as mo
    from oracle import thirdparty as tp
    prob = small_problem
    kep, gaia = _bands(prob)
    lp = LogPosterior(prob, device=0)
    sweep = predictive.ComponentSweep(prob, kep, gaia, device=0, share_grid_with=lp)
    theta = synthetic.sample_prior_box(prob, 300, seed=23, edge_cases=False)
    theta = theta[np.isfinite(lp(theta))][:200]
    assert len(theta) >= 50
    res = predictive.kepler_sweep(sweep, theta, size=64, rng=np.random.default_rng(5))
    rows = res['sample']
    assert rows.shape == (64, 6)

    o = small_oracle.p
    glib = tp.LibraryFilter('Gaia_G', gaia.wavelength, gaia.transmit, 'photon', gaia.vega_zero_flux)
    for k in range(0, 64, 7):
        t = rows[k]
        lg = [mo.get_logg(v, o.matrix) for v in t[:2]]
        out = mo.make_composite_plot(t[:2], lg, t[3:5], t[5], small_oracle.r, o.specs, o.ctm, o.ptm, o.tmi, o.tma,
                                     o.temps, o.lgs, kep, glib)
        _, _, _, _, pm, sm, gp, gs, gm = out
        assert abs(res['kep_contrast'][k] - (sm - pm)) <= 1e-9 * abs(sm - pm) + 1e-12
        assert abs(res['gaia_pri'][k] - gp) <= 1e-9 * abs(gp)
        assert abs(res['gaia_sec'][k] - gs) <= 1e-9 * abs(gs)
        assert abs(res['gaiamag'][k] - gm) <= 1e-9 * abs(gm)
        pc = np.sqrt(1 + 10 ** (-0.4 * (sm - pm)))
        sc = t[4] * np.sqrt(1 + 10 ** (0.4 * (sm - pm)))
        assert abs(res['pri_corr'][k] - pc) <= 1e-9 * pc and abs(res['sec_corr'][k] - sc) <= 1e-9 * sc

    # the sweep handle reuses the fit's grid and does not disturb the fit's own values
    a = lp(theta[:16])
    predictive.write_sweep(str(tmp_path), res)
    for name in ('kep_contrast.txt', 'gaia_sec.txt', 'gaia_pri.txt', 'pri_corr.txt', 'sec_corr.txt'):
        assert np.loadtxt(tmp_path / name).shape == (64,)
    assert np.array_equal(a, lp(theta[:16]))
    sweep.close(); lp.close()
