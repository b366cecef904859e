"""Pins the oracle's restated third-party numerics and helpers against closed forms.

The reference ships no known-answer tests; these are the checks that
stand in for them.  CPU only.
"""
import numpy as np
import pytest

from oracle import mft6_oracle as mo
from oracle import thirdparty as tp


# ---------------------------------------------------------------- extinction (mft6.py:46-64)
def test_ccm89_normalised_at_v():
    # x = 1.82 um^-1 -> y = 0 -> a = 1, b = 0 -> A(lambda) = Av for any Rv
    lam = 1e4 / 1.82
    for rv in (2.1, 3.1, 5.0):
        assert tp.ccm89(lam, 0.73, rv)[0] == pytest.approx(0.73, rel=1e-15)


def test_ccm89_published_ir_ratios():
    # CCM89 Table 3, Rv = 3.1: A_J/A_V = 0.282, A_K/A_V = 0.114 (IR power law, x^1.61)
    assert tp.ccm89(1e4 / 0.80, 1.0)[0] == pytest.approx(0.282, abs=1.5e-3)
    assert tp.ccm89(1e4 / 0.4545, 1.0)[0] == pytest.approx(0.114, abs=1.5e-3)
    x = 0.8
    assert tp.ccm89(1e4 / x, 1.0)[0] == pytest.approx(0.574 * x ** 1.61 - 0.527 * x ** 1.61 / 3.1, rel=1e-15)


def test_ccm89_linear_in_av_and_branch_continuity():
    lam = np.linspace(3100.0, 29000.0, 500)
    assert np.allclose(tp.ccm89(lam, 2.0), 2.0 * tp.ccm89(lam, 1.0), rtol=1e-15)
    eps = 1e-9
    lo = tp.ccm89(1e4 / (1.1 - eps), 1.0)[0]; hi = tp.ccm89(1e4 / (1.1 + eps), 1.0)[0]
    assert abs(lo - hi) < 5e-3                         # CCM89's IR/optical seam is nearly continuous
    with pytest.raises(ValueError):
        tp.ccm89(40000.0, 1.0)                         # x < 0.3


def test_extinct_applies_transmission():
    lam = np.array([5494.505494505494, 12500.0])
    out = mo.extinct(lam, np.array([2.0, 3.0]), 1.0)
    assert out[0] == pytest.approx(2.0 * 10 ** -0.4, rel=1e-14)
    assert out[1] == pytest.approx(3.0 * 10 ** (-0.4 * tp.ccm89(12500.0, 1.0)[0]), rel=1e-15)


# ---------------------------------------------------------------- broadening (mft6.py:124-152)
def test_gaussian_taps_koi2298_window():
    # window 6450-8400 A at 0.2 A, R = 1700 -> 93 taps, sum 1, symmetric
    wl = np.arange(3000, 30000, 0.2)
    w = wl[(wl >= 6450.0) & (wl <= 8400.0)]
    sigma = np.mean(w) / 1700.0 / (2 * np.sqrt(2 * np.log(2)))
    nx, e = tp.gauss_kernel_taps(w[1] - w[0], sigma, 5)
    assert len(e) == 93 and len(w) == 9750
    assert e.sum() == pytest.approx(1.0, abs=1e-15)
    assert np.allclose(e, e[::-1], rtol=1e-13) and np.allclose(nx, -nx[::-1], atol=1e-12)


def test_broaden_constant_and_edge_patch():
    wl = 6000.0 + 0.2 * np.arange(3000)
    _, b = mo.broaden(wl, np.full(wl.size, 7.0), 1700)
    assert np.allclose(b[60:-60], 7.0, rtol=1e-14)            # interior untouched
    assert np.all(b[:5] == b[5]) and np.all(b[-10:] == b[-11])  # mft6.py:129-130
    assert b[5] < 7.0                                         # zero-padding droop is kept, not fixed


def test_broaden_preserves_line_position_and_flux():
    wl = 6000.0 + 0.2 * np.arange(4000)
    line = np.exp(-0.5 * ((wl - 6400.0) / 0.5) ** 2)
    _, b = mo.broaden(wl, line, 1700)
    assert abs(wl[np.argmax(b)] - 6400.0) < 0.11
    assert b.sum() == pytest.approx(line.sum(), rel=1e-12)


def test_instr_broad_requires_equidistant():
    with pytest.raises(ValueError):
        tp.instr_broad_gauss_fast(np.array([1.0, 2.0, 4.0, 5.0]), np.ones(4), 100)


# ---------------------------------------------------------------- pyphot get_flux (mft6.py:776)
def test_flat_spectrum_through_any_filter_is_flat():
    lam = np.linspace(5000, 9000, 4001)
    f = tp.LibraryFilter('x', np.linspace(6000, 7000, 50), np.hanning(50), 'photon')
    assert f.get_flux(lam, np.full(lam.size, 3.3e-15)) == pytest.approx(3.3e-15, rel=1e-13)
    g = tp.LibraryFilter('y', np.linspace(6000, 7000, 50), np.hanning(50), 'energy')
    assert g.get_flux(lam, np.full(lam.size, 3.3e-15)) == pytest.approx(3.3e-15, rel=1e-13)


def test_photon_vs_energy_weighting_on_a_slope():
    lam = np.linspace(5000, 9000, 40001)
    T = np.where((lam >= 6000) & (lam <= 7000), 1.0, 0.0)
    fp = tp.LibraryFilter('p', lam, T, 'photon'); fe = tp.LibraryFilter('e', lam, T, 'energy')
    # f = lam: energy -> <lam> = 6500 ; photon -> <lam^2>/<lam> = (7000^3-6000^3)/3 / ((7000^2-6000^2)/2)
    assert fe.get_flux(lam, lam) == pytest.approx(6500.0, rel=1e-4)
    assert fp.get_flux(lam, lam) == pytest.approx((7000 ** 3 - 6000 ** 3) / 3 / ((7000 ** 2 - 6000 ** 2) / 2), rel=1e-4)


def test_ab_zero_flux_definition():
    f = tp.LibraryFilter('b', np.linspace(6000, 7000, 200), np.ones(200), 'photon')
    lp = f.lpivot
    assert lp == pytest.approx(np.sqrt((7000 ** 2 - 6000 ** 2) / 2 / np.log(7000 / 6000)), rel=1e-6)
    # AB: f_nu = 10^(-48.6/2.5) erg/s/cm^2/Hz -> f_lambda = f_nu * c / lpivot^2
    assert f.AB_zero_flux == pytest.approx(10 ** (-0.4 * 48.6) * tp.C_AA_PER_S / lp ** 2, rel=1e-12)


# ---------------------------------------------------------------- bayestar restatement
def test_bayestar_sightline_distance_handling():
    edges = np.array([4.0, 4.5, 5.0, 5.5])
    samples = np.array([[0.1, 0.2, 0.4, 0.8], [0.0, 0.1, 0.1, 0.3]])
    s = tp.BayestarSightline(edges, samples)
    d = lambda dm: 10 ** (dm / 5 + 1)                                   # noqa: E731  pc
    assert np.allclose(s.query(d(4.75)), [0.3, 0.1])                    # linear in DM
    assert np.allclose(s.query(d(9.0)), [0.8, 0.3])                     # beyond the last slice
    assert np.allclose(s.query(d(3.0)), 10 ** (0.2 * (3.0 - 4.0)) * samples[:, 0])   # nearer than the first
    assert np.allclose(s.query(d(4.5)), [0.2, 0.1])                     # on an edge


# ---------------------------------------------------------------- mft6 helpers
def test_find_nearest_ties_go_low():
    assert mo.find_nearest([3000, 3100, 3200], 3050.0) == 0
    assert mo.find_nearest([3000, 3100, 3200], 3051.0) == 1


def test_interp_2_spec_either_order():
    a, b = np.array([1.0, 2.0]), np.array([3.0, 6.0])
    assert np.allclose(mo.interp_2_spec(a, b, 3000, 3100, 3025.0), a + 0.25 * (b - a))
    assert np.allclose(mo.interp_2_spec(b, a, 3100, 3000, 3025.0), a + 0.25 * (b - a))   # nearest node above


def test_chisq_uses_error_squared():
    assert np.allclose(mo.chisq([3.0], [1.0], [-0.5]), [16.0])          # sign of the "error" is irrelevant


def test_norm_spec_exact_quadratic():
    wl = np.linspace(0.645, 0.84, 500)
    model = 1.0 + 0.1 * np.sin(40 * wl)
    quad = 0.9 + 0.8 * (wl - 0.7) - 3.0 * (wl - 0.7) ** 2
    out = mo.norm_spec(wl, model, model * quad)
    assert np.allclose(out, model, rtol=1e-12)


def _toy_specs():
    wl = np.linspace(5000.0, 9000.0, 401)
    temps, lgs = [3000, 3100, 3200], [4.5, 5.0]
    specs = {'wl': wl}
    for t in temps:
        for g in lgs:
            specs['{}, {}'.format(t, g)] = (t / 1000.0) * wl + 100.0 * g
    return specs, temps, lgs, wl


def test_get_spec_nodes_and_bilinear():
    specs, temps, lgs, wl = _toy_specs()
    w, f = mo.get_spec(3100.0, 5.0, [0.55, 0.85], specs, temps, lgs)
    keep = (wl >= 5500.0) & (wl <= 8500.0)
    assert np.array_equal(w, wl[keep]) and np.array_equal(f, specs['3100, 5.0'][keep])      # on a node: the row itself
    # off-node: the toy grid is bilinear in (T, g), so interpolation is exact
    w, f = mo.get_spec(3140.0, 4.8, [0.55, 0.85], specs, temps, lgs)
    assert np.allclose(f, 3.14 * wl[keep] + 480.0, rtol=1e-14)
    w, f = mo.get_spec(3100.0, 4.6, [0.55, 0.85], specs, temps, lgs)                        # on-node in T only
    assert np.allclose(f, 3.1 * wl[keep] + 460.0, rtol=1e-14)
    with pytest.raises(IndexError):
        mo.get_spec(3100.0, 5.2, [0.55, 0.85], specs, temps, lgs)                           # above the log g grid


def test_isochrone_cut_and_interpolation():
    from mcmc_spec_b200 import synthetic
    m = synthetic.make_isochrone_matrix()
    track = m[m[:, 1] == 9.0]
    assert track.shape[0] > 220
    t = 0.5 * (track[10, 4] + track[11, 4])
    assert float(mo.get_logg(t, m)) == pytest.approx(0.5 * (track[10, 5] + track[11, 5]), rel=1e-13)
    assert float(mo.get_logg(track[7, 4], m)) == pytest.approx(track[7, 5], rel=1e-15)
    lum = 0.5 * (track[10, 6] + track[11, 6])
    r = np.sqrt(lum * mo.LSUN / (4 * np.pi * mo.SIGMA_SB * t ** 4)) / mo.RSUN
    assert float(mo.get_radius(t, m)) == pytest.approx(r, rel=1e-13)
    with pytest.raises(ValueError):
        mo.get_logg(2000.0, m)                                                               # interp1d bounds_error


def test_logprior_bounds_and_terms():
    from mcmc_spec_b200 import synthetic
    m = synthetic.make_isochrone_matrix()
    dust = tp.BayestarSightline(*synthetic.make_dust_table())
    prior = [0] * 10 + [2e-3, 1e-4]
    ok = [3850.0, 3425.0, 0.1, 0.45, 0.6, 2e-3]
    kw = dict(prior=prior, ext=True, dist_fit=True, rad_prior=False)
    base = mo.logprior(ok, 2, 3000, 5500, m, dust, **kw)
    assert np.isfinite(base)
    for bad in ([5501.0, 3425, 0.1, 0.45, 0.6, 2e-3], [3850, 2999.0, 0.1, 0.45, 0.6, 2e-3],
                [3850, 3425, -1e-9, 0.45, 0.6, 2e-3], [3850, 3425, 0.1, 0.049, 0.6, 2e-3],
                [3850, 3425, 0.1, 1.51, 0.6, 2e-3], [3850, 3425, 0.1, 0.45, 0.049, 2e-3],
                [3850, 3425, 0.1, 0.45, 0.6, 1 / 3001.0], [3850, 3425, 0.1, 0.45, 0.6, 0.2501]):
        assert mo.logprior(bad, 2, 3000, 5500, m, dust, **kw) == -np.inf
    # parallax Gaussian is the only user prior (mft6.py:3689)
    shifted = list(ok); shifted[5] = 2e-3 + 1e-4
    av = dust.query(1 / 2e-3) * 3.1 * 0.884
    av2 = dust.query(1 / shifted[5]) * 3.1 * 0.884
    def avt(a):                                                                             # noqa: E306
        s = np.std(a) if np.std(a) != 0 else 0.05
        return -0.5 * ((0.1 - np.mean(a)) / s) ** 2
    assert mo.logprior(shifted, 2, 3000, 5500, m, dust, **kw) - avt(av2) == pytest.approx(base - avt(av) - 0.5, rel=1e-12)
    # radius prior: zero penalty exactly on the isochrone
    r1, r2 = float(mo.get_radius(3850.0, m)), float(mo.get_radius(3425.0, m))
    on = [3850.0, 3425.0, 0.1, r1, r2 / r1, 2e-3]
    a = mo.logprior(on, 2, 3000, 5500, m, dust, prior=prior, ext=True, dist_fit=True, rad_prior=True)
    b = mo.logprior(on, 2, 3000, 5500, m, dust, prior=prior, ext=True, dist_fit=True, rad_prior=False)
    assert a == pytest.approx(b, abs=1e-20)


def test_splice_broadened_grid_layout():
    wl = np.arange(5000.0, 7000.0, 0.2)
    native = {'3000, 4.5': 1.0 + 0.1 * np.sin(wl / 0.3)}      # period 1.9 A << FWHM 3.5 A
    specs = mo.splice_broadened_grid(native, wl, [5600.0, 6400.0], 1700)
    assert np.array_equal(specs['wl'], wl)
    out = specs['3000, 4.5']
    inw = (wl >= 5600.0) & (wl <= 6400.0)
    assert np.array_equal(out[~inw], native['3000, 4.5'][~inw])           # outside the window: native resolution
    assert np.std(out[inw][100:-100]) < 0.5 * np.std(native['3000, 4.5'][inw][100:-100])


def test_median_even_odd_is_numpy():
    # the CUDA path's median is compared against np.median through the oracle; record the convention
    assert np.median([3.0, 1.0, 2.0, 10.0]) == 2.5 and np.median([3.0, 1.0, 2.0]) == 2.0


def test_plot_branch_is_consistent_with_the_contrast_integral(small_problem, small_oracle):
    """make_composite(plot=True) integrates each component through the Kepler curve exactly as the
    contrast loop does (mft6.py:798 vs :729-731): feeding a contrast filter's curve must return that
    filter's contrast; the composite Gaia flux is the sum of the component fluxes."""
    from oracle import mft6_oracle as mo
    from oracle import thirdparty as tp
    o = small_oracle.p
    t = small_problem.truth
    lg = [mo.get_logg(v, o.matrix) for v in t[:2]]
    w, spec, con, cwl, ph = mo.make_composite(t[:2], lg, t[3:5], t[5], o.fr[2], o.fr[5], small_oracle.r, o.specs,
                                              o.ctm, o.ptm, o.tmi, o.tma, o.temps, o.lgs, o.photlib, fs=o.fs)
    kep = (o.ctm[0][0], o.ctm[1][0])
    b = small_problem.photbands[0]
    glib = tp.LibraryFilter('Gaia_G', b.wavelength, b.transmit, 'photon', 3.0e-9)
    out = mo.make_composite_plot(t[:2], lg, t[3:5], t[5], small_oracle.r, o.specs, o.ctm, o.ptm, o.tmi, o.tma,
                                 o.temps, o.lgs, kep, glib)
    wl, spec1, pri, sec, pm, sm, gp, gs, gm = out
    assert np.array_equal(wl, w) and np.allclose(spec1, spec, rtol=1e-15)
    assert abs((sm - pm) - con[0]) < 1e-12
    fl = 10 ** (-0.4 * gp) + 10 ** (-0.4 * gs)
    assert abs(-2.5 * np.log10(fl) - gm) < 1e-12


def test_correction_factors_closed_form(tmp_path):
    """mft6.py:2531-2532: equal-brightness components -> sqrt(2) and ratio*sqrt(2); files are np.savetxt columns."""
    from mcmc_spec_b200 import predictive
    pc, sc = predictive.correction_factors(np.array([0.0, 2.5]), np.array([1.0, 0.5]))
    assert np.allclose(pc, [np.sqrt(2.0), np.sqrt(1.1)]) and np.allclose(sc, [np.sqrt(2.0), 0.5 * np.sqrt(11.0)])
    res = dict(kep_contrast=[0.0, 2.5], gaia_sec=[1.0, 2.0], gaia_pri=[3.0, 4.0], pri_corr=pc, sec_corr=sc)
    predictive.write_sweep(str(tmp_path), res)
    assert np.allclose(np.loadtxt(tmp_path / 'sec_corr.txt'), sc)
    assert predictive.percentile_summary([-1.0, -2.0, -3.0])[0] == 2.0


def test_linearity_identities_behind_the_bracket_only_mode(small_problem, small_oracle):
    """What ms_set_fast relies on (algebraic note), checked on the oracle itself:
    (i) a component's band flux equals the interpolation-weighted sum of the band fluxes of its bracketing grid
        rows (np.trapz and the grid interpolation are linear, the weights do not depend on wavelength);
    (ii) the rebinned model at an observed pixel depends only on the two model pixels that bracket it."""
    from oracle import mft6_oracle as mo
    o = small_oracle.p
    t = np.array(small_problem.truth, dtype=float)
    t[0], t[1] = 3733.0, 3541.0                                     # off-node in Teff; log g off-node via the isochrone
    lg = [float(mo.get_logg(v, o.matrix)) for v in t[:2]]
    region_um = [min(small_oracle.r) - 0.2, max(small_oracle.r) + 0.2]
    wl, spec = mo.get_spec(t[0], lg[0], region_um, o.specs, o.temps, o.lgs)
    # (i) band flux of the interpolated spectrum vs the bilinear combination of node band fluxes
    ran, tm = np.array(o.ctm[0][0]), np.array(o.ctm[1][0])
    inb = (wl >= ran.min()) & (wl <= ran.max())
    tran = np.interp(wl[inb], ran, tm)
    direct = np.trapezoid(spec[inb] * tran, wl[inb])
    t1, t2 = mo._bracket(o.temps, t[0]); g1, g2 = mo._bracket(o.lgs, lg[0])
    wt, wg = (t[0] - t1) / (t2 - t1), (lg[0] - g1) / (g2 - g1)
    full = np.asarray(o.specs['wl'])
    keep = (full >= wl[0]) & (full <= wl[-1])
    node_flux = {k: np.trapezoid(np.asarray(o.specs['{}, {}'.format(*k)])[keep][inb] * tran, wl[inb])
                 for k in ((t1, g1), (t1, g2), (t2, g1), (t2, g2))}
    combo = ((1 - wt) * ((1 - wg) * node_flux[(t1, g1)] + wg * node_flux[(t1, g2)])
             + wt * ((1 - wg) * node_flux[(t2, g1)] + wg * node_flux[(t2, g2)]))
    assert abs(combo - direct) <= 1e-13 * abs(direct)
    # (ii) linear rebinning reads two model pixels per observed pixel
    x_new = np.asarray(small_problem.data[0]) * 1e4
    ext = mo.extinct(wl, spec, 0.3)
    hi = np.searchsorted(wl, x_new, side='right'); lo = hi - 1
    two = ext[lo] + (ext[hi] - ext[lo]) * (x_new - wl[lo]) / (wl[hi] - wl[lo])
    from scipy.interpolate import interp1d
    assert np.allclose(two, interp1d(wl, ext)(x_new), rtol=1e-14, atol=0)
