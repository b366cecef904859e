"""Seeded synthetic inputs of the reference's shapes (the real model grid is not available offline).

The BT-Settl spectra (mft6.py:324), the MIST isochrone table (mft6.py:3487) and the Bayestar map
(mft6.py:23) are not in the reference repository, so every table they would provide is generated
here from a seed (default 2298), in float64, with the structure the hot path relies on:

  * node spectra  pi*B_lambda(T) * (1 - sum of Gaussian lines whose depths vary smoothly with
    (T, g[, z])) * (1 + 0.02*eta_node(lambda)), erg/s/cm^2/A at the stellar surface, on the
    reference's common grid ``arange(specmin, specmax, dlam)`` (mft6.py:343);
  * the data window of every row broadened to ``res`` and spliced back, i.e. the operator of
    ``spec_interpolator`` (mft6.py:366-383): sum-normalised sampled Gaussian, zero-padded 'same'
    convolution, first 5 / last 10 samples patched (mft6.py:128-130);
  * a MIST-like table whose log-age == 9.0 rows carry a monotone Teff track with log g and L;
  * a Bayestar-like sightline: 120 distance-modulus slices x 5 samples, non-decreasing;
  * passbands either read from a ``bps/`` directory or synthetic flat-topped curves;
  * an observed spectrum drawn from the model at a seeded truth.

This module produces *inputs*.  It is not the product's likelihood path and it is not the
oracle: the forward model used to draw the observed spectrum below only has to be plausible.
"""
from __future__ import annotations

import numpy as np

from .passbands import PhotBand, ab_zero_flux, band_flux
from .problem import Problem

H = 6.62607015e-27
C = 2.99792458e10
KB = 1.380649e-16
SIGMA_SB = 5.670374e-5
LSUN = 3.839e33
RSUN = 6.957e10
PC_CM = 3.086e18


# ------------------------------------------------------------------------------ spectral grid

def planck_surface_flux(wl_aa, teff):
    """pi * B_lambda(T) in erg/s/cm^2/A."""
    lam = wl_aa * 1e-8
    x = H * C / (lam * KB * teff)
    return np.pi * 2.0 * H * C * C / lam ** 5 / np.expm1(x) * 1e-8


def gauss_taps(dx, mean_wl, res, maxsig=5):
    """Tap offsets and sum-normalised weights of the instrumental Gaussian (mft6.py:128)."""
    fwhm = mean_wl / float(res)
    sigma = fwhm / (2.0 * np.sqrt(2.0 * np.log(2.0)))
    lx = int(((sigma * maxsig) / dx) * 2.0) + 1
    nx = (np.arange(lx, dtype=int) - sum(divmod(lx, 2)) + 1) * dx
    e = np.exp(-nx ** 2 / (2.0 * sigma ** 2))
    return nx, e / e.sum()


def broaden_window_rows(rows, wl, window, res):
    """Apply the ``spec_interpolator`` window-broadening splice to every row (numpy, in place)."""
    inwin = np.where((wl >= min(window)) & (wl <= max(window)))[0]
    if inwin.size < 16:
        return rows
    w = wl[inwin]
    _, taps = gauss_taps(w[1] - w[0], np.mean(w), res)
    lo, hi = inwin[0], inwin[-1] + 1
    flat = rows.reshape(-1, rows.shape[-1])
    for row in flat:
        b = np.convolve(row[lo:hi], taps, mode='same')
        b[0:5] = b[5]
        b[len(b) - 10:] = b[len(b) - 11]
        row[lo:hi] = b
    return rows


def make_native_grid(temps, lgs, wl, seed=2298, fehs=None, n_lines=400):
    """Un-broadened synthetic node spectra, shape [n_logg, n_teff, P] (or with a leading [Fe/H])."""
    rng = np.random.default_rng(seed)
    P = wl.shape[0]
    centers = rng.uniform(wl[0], wl[-1], n_lines)
    widths = rng.uniform(0.3, 3.0, n_lines)
    depth0 = rng.uniform(0.0, 0.6, n_lines)
    tper = rng.uniform(600.0, 2500.0, n_lines)
    tphase = rng.uniform(0, 2 * np.pi, n_lines)
    gsens = rng.uniform(-1.0, 1.0, n_lines)
    zsens = rng.uniform(-1.0, 1.0, n_lines)
    dlam = (wl[-1] - wl[0]) / (P - 1)
    # line profiles on their local pixel windows (node independent)
    profiles = []
    for j in range(n_lines):
        half = int(np.ceil(6.0 * max(widths[j], dlam) / dlam))
        c = int(round((centers[j] - wl[0]) / dlam))
        lo, hi = max(0, c - half), min(P, c + half + 1)
        if hi <= lo:
            profiles.append((0, 0, None))
            continue
        prof = np.exp(-0.5 * ((wl[lo:hi] - centers[j]) / max(widths[j], 0.6 * dlam)) ** 2)
        profiles.append((lo, hi, prof))
    eta_period = np.array([53.0, 171.0, 409.0, 977.0, 2903.0])
    eta_amp = np.array([0.30, 0.25, 0.20, 0.15, 0.10])

    zs = [0.0] if fehs is None else list(fehs)
    out = np.empty((len(zs), len(lgs), len(temps), P), dtype=np.float64)
    for iz, z in enumerate(zs):
        for ig, g in enumerate(lgs):
            for it, t in enumerate(temps):
                depth = depth0 * (0.55 + 0.45 * np.sin(t / tper * 2 * np.pi + tphase))
                depth = depth * (1.0 + 0.15 * (g - 4.75) * gsens) * (1.0 + 0.3 * z * zsens)
                depth = np.clip(depth, 0.0, 0.7)
                absorb = np.zeros(P)
                for j, (lo, hi, prof) in enumerate(profiles):
                    if prof is not None:
                        absorb[lo:hi] += depth[j] * prof
                phases = rng.uniform(0, 2 * np.pi, eta_period.size)
                eta = (eta_amp[:, None] * np.sin(2 * np.pi * wl[None, :] / eta_period[:, None]
                                                 + phases[:, None])).sum(0)
                out[iz, ig, it] = (planck_surface_flux(wl, float(t)) * (1.0 - np.minimum(absorb, 0.9))
                                   * (1.0 + 0.02 * eta) * (1.0 + 0.05 * (g - 4.75)) * (1.0 - 0.08 * z))
    return out[0] if fehs is None else out


# ------------------------------------------------------------------------------ small tables

def make_isochrone_matrix(seed=2298, n_track=260):
    """MIST-like table: col 1 log age, col 4 Teff [K], col 5 log g, col 6 L/Lsun (already linear).

    The log-age == 9.0 rows form the track; the first 220 rise monotonically from 2800 K to
    6500 K (log g falls from ~5.2 to ~4.85, crossing the 5.0 node), the rest turn back to cooler
    temperatures so that forgetting the ``[:220]`` cut of mft6.py:76/:95 changes the answer.
    Rows of other ages are interleaved so the age selection matters too.
    """
    rng = np.random.default_rng(seed + 1)
    t_main = 2800.0 + 3700.0 * (np.arange(220) / 219.0) ** 1.15
    t_back = np.linspace(6480.0, 5200.0, n_track - 220)
    teff = np.concatenate([t_main, t_back])
    frac = (teff - 2800.0) / 3700.0
    logg = 5.2 - 0.35 * frac + 0.004 * np.sin(frac * 23.0 + rng.uniform(0, 6.28))
    logg[220:] -= 0.4
    radius = 0.13 + 0.95 * np.clip(frac, 0, None) ** 1.1 + 0.003 * np.sin(frac * 17.0)
    radius[220:] *= 1.6
    lum = 4 * np.pi * (radius * RSUN) ** 2 * SIGMA_SB * teff ** 4 / LSUN
    rows = []
    for k in range(n_track):
        for age in (8.95, 9.0, 9.05):
            scale = 1.0 + 0.01 * (age - 9.0) * 20
            rows.append([0.0, age, 0.1 + 0.005 * k, 0.0, teff[k] * scale, logg[k] / scale, lum[k] * scale,
                         0.0])
    return np.array(rows, dtype=np.float64)


def make_dust_table(seed=2298, n_dm=120, n_samples=5):
    """Bayestar2019-like sightline: DM edges 4..19 step 0.125, cumulative reddening samples."""
    rng = np.random.default_rng(seed + 2)
    dm_edges = 4.0 + 0.125 * np.arange(n_dm)
    incr = rng.gamma(0.35, 0.004, size=(n_samples, n_dm))
    incr[:, rng.random(n_dm) < 0.5] = 0.0          # flat stretches: exercises sigma handling
    samples = 0.004 + np.cumsum(incr, axis=1)
    return dm_edges, samples


def synthetic_passband(center, width, n=400, edge=0.12, peak=0.9, ripple_seed=0):
    """Flat-topped curve with raised-cosine edges and a small ripple; (wl [A], T)."""
    rng = np.random.default_rng(1000 + ripple_seed)
    wl = np.linspace(center - 0.5 * width, center + 0.5 * width, n)
    u = (wl - wl[0]) / (wl[-1] - wl[0])
    ramp = np.clip(np.minimum(u, 1 - u) / edge, 0, 1)
    T = peak * 0.5 * (1 - np.cos(np.pi * ramp)) * (1 + 0.03 * np.sin(u * rng.uniform(20, 60)))
    T[0] = T[-1] = 0.0
    return wl, T


def _filter_table(curves, res):
    """[[wl lists], [T lists], [n_resel], [mean wl]] like mft6.py:3579-3595."""
    wls, tras, nre, cwl = [], [], [], []
    for w, t in curves:
        wls.append(list(w)); tras.append(list(t))
        nre.append((max(w) - min(w)) / (np.mean(w) / res)); cwl.append(np.mean(w))
    return [wls, tras, nre, cwl]


# ------------------------------------------------------------------------------ forward model
# (data generation only: draws an observed spectrum / photometry consistent with the tables)

def _ccm89_k(wl_aa):
    x = 1.0e4 / np.asarray(wl_aa, dtype=np.float64)
    y = x - 1.82
    a_o = 1 + y * (0.17699 + y * (-0.50447 + y * (-0.02427 + y * (0.72085 + y * (0.01979 + y * (-0.77530 + y * 0.32999))))))
    b_o = y * (1.41338 + y * (2.28305 + y * (1.07233 + y * (-5.38434 + y * (-0.62251 + y * (5.30260 - y * 2.09002))))))
    a = np.where(x < 1.1, 0.574 * x ** 1.61, a_o)
    b = np.where(x < 1.1, -0.527 * x ** 1.61, b_o)
    return a + b / 3.1


def _interp_node_axis(nodes, v):
    nodes = np.asarray(nodes, dtype=np.float64)
    j = int(np.clip(np.searchsorted(nodes, v, side='right') - 1, 0, len(nodes) - 2))
    u = (v - nodes[j]) / (nodes[j + 1] - nodes[j])
    return j, u


def _draw_model(prob_like, theta):
    grid, wl, temps, lgs, matrix = (prob_like['grid'], prob_like['wl'], prob_like['temps'],
                                    prob_like['lgs'], prob_like['matrix'])
    ns = (len(theta) - 2) // 2
    av, r1, plx = theta[ns], theta[ns + 1], theta[-1]
    stars = [(theta[0], r1)] + [(theta[k], r1 * theta[ns + 1 + k]) for k in range(1, ns)]
    track = matrix[matrix[:, 1] == 9.0][:220]
    comps = []
    for t, rad in stars:
        g = np.interp(t, track[:, 4], track[:, 5])
        jt, ut = _interp_node_axis(temps, t)
        jg, ug = _interp_node_axis(lgs, g)
        g2 = grid if grid.ndim == 3 else prob_like['grid_z']
        s = ((1 - ug) * ((1 - ut) * g2[jg, jt] + ut * g2[jg, jt + 1])
             + ug * ((1 - ut) * g2[jg + 1, jt] + ut * g2[jg + 1, jt + 1]))
        comps.append(s * (rad * RSUN / (PC_CM / plx)) ** 2)
    return comps


def isochrone_radius(matrix, teff):
    """Radius [Rsun] the isochrone assigns to ``teff`` (data generation; the relation of mft6.py:66-85)."""
    track = matrix[matrix[:, 1] == 9.0][:220]
    o = np.argsort(track[:, 4], kind='mergesort')
    lum = np.interp(teff, track[o, 4], track[o, 6])
    return np.sqrt(lum * LSUN / (4 * np.pi * SIGMA_SB * teff ** 4)) / RSUN


def consistent_truth(t1, t2, av, plx, matrix=None, seed=2298):
    """A binary whose radii sit on the isochrone, so that the 2 % radius prior (mft6.py:1262-1269) and the
    likelihood agree and a fit can recover the truth: [T1, T2, Av, R(T1), R(T2)/R(T1), plx]."""
    matrix = make_isochrone_matrix(seed) if matrix is None else matrix
    r1, r2 = isochrone_radius(matrix, float(t1)), isochrone_radius(matrix, float(t2))
    return np.array([t1, t2, av, r1, r2 / r1, plx], dtype=np.float64)


def slice_feh(grid4, fehs, feh):
    """Linear [Fe/H] interpolation of a 4-D grid to the fit's fixed metallicity (data generation)."""
    j, u = _interp_node_axis(fehs, feh)
    return (1 - u) * grid4[j] + u * grid4[j + 1]


# ------------------------------------------------------------------------------ problem builders

def make_problem(*, temps=None, lgs=None, specmin=3000.0, specmax=30000.0, dlam=0.2, n_pix=None,
                 window_um=(0.645, 0.84), res=1700, n_data=1945, data_wl_um=None,
                 contrast_curves=None, contrast_names=None, phot_bands=None, phot_curves=None,
                 phot_names=None, truth=None, plx=2.0732e-3, plx_err=0.0277e-3, snr=150.0,
                 seed=2298, fehs=None, feh=0.0, rad_prior=True, dist_fit=True,
                 data_flux=None, data_err=None, cmag=None, cerr=None, pmag=None, perr=None,
                 ctm=None, ptm=None, pfilt_names=None, nspec=2, nospec=False, feh_free=False):
    """Assemble a synthetic ``Problem``.

    Defaults give the reference's koi2298 shapes (26 x 4 nodes, 135 000-pixel rows at 0.2 A,
    window 6450-8400 A at R = 1700) with synthetic passbands; pass ``contrast_curves`` /
    ``phot_bands`` read from ``bps/`` (see ``passbands``) for the real filter set, and
    ``data_wl_um`` / ``data_flux`` / ``data_err`` for a real observed spectrum.  ``ctm`` / ``ptm``
    given explicitly (as ``passbands.assemble_filter_tables`` returns them) override the tables
    derived from the curves, so reference quirks such as float32 FITS columns survive.
    """
    rng = np.random.default_rng(seed + 7)
    temps = list(range(3000, 5501, 100)) if temps is None else [int(t) for t in temps]
    lgs = [4.0, 4.5, 5.0, 5.5] if lgs is None else [float(g) for g in lgs]
    if n_pix is None:
        wl = np.arange(specmin, specmax, dlam)          # mft6.py:343
    else:
        wl = specmin + dlam * np.arange(n_pix, dtype=np.float64)

    native = make_native_grid(temps, lgs, wl, seed=seed, fehs=fehs)
    window_aa = [window_um[0] * 1e4, window_um[1] * 1e4]
    grid = broaden_window_rows(native, wl, window_aa, res)

    matrix = make_isochrone_matrix(seed)
    dm_edges, dust_samples = make_dust_table(seed)

    # ---- passbands
    if contrast_curves is None:
        span = wl[-1] - wl[0]
        contrast_curves = [synthetic_passband(wl[0] + 0.20 * span, 0.16 * span, 900, ripple_seed=1),
                           synthetic_passband(wl[0] + 0.86 * span, 0.18 * span, 120, ripple_seed=2)]
        contrast_names = ['synth_c0', 'synth_c1']
    if phot_bands is None:
        if phot_curves is None:
            span = wl[-1] - wl[0]
            fr_c = [0.12, 0.26, 0.40, 0.55, 0.70, 0.88]
            fr_w = [0.07, 0.08, 0.10, 0.09, 0.12, 0.13]
            phot_curves = [synthetic_passband(wl[0] + c * span, w * span, 300 + 150 * k, ripple_seed=10 + k)
                           for k, (c, w) in enumerate(zip(fr_c, fr_w))]
            phot_names = ['SDSS_r', 'SDSS_i', 'SDSS_z', '2MASS_J', '2MASS_H', '2MASS_Ks']
        phot_bands = []
        for k, ((w, t), name) in enumerate(zip(phot_curves, phot_names)):
            system = 'vega' if ('2MASS' in name or k % 2) and 'SDSS' not in name else 'ab'
            band = PhotBand(name, np.asarray(w, float), np.asarray(t, float), 'photon', system)
            band.ab_zero_flux = float(ab_zero_flux(band))
            # synthetic "Vega": a 9600 K blackbody diluted to ~Vega's flux level
            band.vega_zero_flux = float(band_flux(band, band.wavelength,
                                                  planck_surface_flux(band.wavelength, 9600.0) * 6.2e-17))
            phot_bands.append(band)
    if contrast_names is None:
        contrast_names = [f'c{k}' for k in range(len(contrast_curves))]
    if ctm is None:
        ctm = _filter_table(contrast_curves, res)
    else:
        contrast_curves = list(zip(ctm[0], ctm[1]))
    # ptm: the reference carries the bps/ curves of the photometric filters only for the region
    # span and the central wavelengths (mft6.py:3588-3595); here they coincide with the library.
    if ptm is None:
        ptm = _filter_table([(b.wavelength, b.transmit) for b in phot_bands], res)
    tmi = min([min(w) for w in ctm[0]] + [min(w) for w in ptm[0]])          # mft6.py:3577-3597 (either list may be empty)
    tma = max([max(w) for w in ctm[0]] + [max(w) for w in ptm[0]])

    # ---- truth and observed data
    if truth is None:
        truth = (np.array([3850.0, 3425.0, 0.12, 0.4994, 0.61, plx]) if nspec == 2 else
                 np.array([3850.0, 3600.0, 3300.0, 0.12, 0.4994, 0.78, 0.5, plx]))
    truth = np.asarray(truth, dtype=np.float64)
    like = dict(grid=grid if fehs is None else None, wl=wl, temps=temps, lgs=lgs, matrix=matrix)
    if fehs is not None:
        like['grid'] = slice_feh(grid, fehs, feh)
    comps = _draw_model(like, truth)
    pri, sec = comps[0], comps[1]
    comp = sum(comps)
    if data_wl_um is None:
        lo, hi = window_um
        step = (hi - lo) / (n_data + 1)
        data_wl_um = lo + step * (np.arange(n_data) + 1) + rng.uniform(-0.05, 0.05, n_data) * step
        data_wl_um = np.sort(data_wl_um)
    data_wl_um = np.asarray(data_wl_um, dtype=np.float64)
    if data_flux is None:
        ext = comp * 10.0 ** (-0.4 * truth[nspec] * _ccm89_k(wl))
        clean = np.interp(data_wl_um * 1e4, wl, ext)
        av_true = truth[nspec]
        cont = 1.0 + 0.03 * (data_wl_um - data_wl_um.mean()) / np.ptp(data_wl_um)
        sigma = clean / snr
        flux = clean * cont + rng.normal(0.0, sigma)
        med = np.median(flux)
        data_err = sigma / med
        data_flux = flux / med
    data_flux = np.asarray(data_flux, dtype=np.float64)
    data_err = np.asarray(data_err, dtype=np.float64)

    # ---- observed contrasts / photometry from the truth (plus seeded offsets)
    if cmag is None:
        cmag, cerr = [], []
        for w, t in contrast_curves:
            w = np.asarray(w, float); t = np.asarray(t, float)
            o = np.argsort(w, kind='mergesort')
            inb = (wl >= w.min()) & (wl <= w.max())
            tr = np.interp(wl[inb], w[o], t[o])
            m = [-2.5 * np.log10(np.trapezoid(s[inb] * tr, wl[inb])) for s in comps]
            other = 1 if (nspec == 2 or len(cmag) < len(contrast_curves) // 2) else 2          # mft6.py:747-749
            cerr.append(0.05)
            cmag.append(float(m[other] - m[0] + rng.normal(0, 0.03)))
    if pmag is None:
        pmag, perr = [], []
        for b in phot_bands:
            m = -2.5 * np.log10(band_flux(b, wl, comp) / b.zero_flux) + truth[nspec] * float(_ccm89_k(np.mean(b.wavelength)))
            perr.append(0.03)
            pmag.append(float(m + rng.normal(0, 0.02)))
    phot_names_fr = list(pfilt_names) if pfilt_names is not None else [b.name for b in phot_bands]
    fr = [list(cmag), list(cerr), np.array(contrast_names), np.array(pmag, dtype=float), list(perr),
          np.array(phot_names_fr)]
    ndim = 2 * nspec + 2
    prior = [*np.zeros(2 * ndim - 2), plx, plx_err]              # mft6.py:3689
    return Problem(grid=grid, wl=wl, temps=temps, lgs=lgs, data=[data_wl_um, data_flux], err=data_err,
                   fr=fr, ctm=ctm, ptm=ptm, tmi=tmi, tma=tma, matrix=matrix, dust_dm_edges=dm_edges,
                   dust_samples=dust_samples, photbands=phot_bands, prior=prior, nspec=nspec, nospec=nospec, av=True,
                   dist_fit=dist_fit, rad_prior=rad_prior, fehs=None if fehs is None else list(fehs),
                   feh=feh, feh_free=bool(feh_free), truth=truth, meta=dict(seed=seed, res=res, window_um=tuple(window_um)))


def make_problem_small(seed=2298, **kw):
    """A reduced problem (fast on the CPU oracle): 6 x 4 nodes, 6000-pixel rows, 300 data pixels."""
    args = dict(temps=range(3400, 3901, 100), lgs=[4.0, 4.5, 5.0, 5.5], specmin=5400.0, dlam=0.5,
                n_pix=6000, window_um=(0.62, 0.74), n_data=300, res=1700, seed=seed,
                truth=[3720.0, 3530.0, 0.10, 0.45, 0.7, 2.5e-3], plx=2.5e-3, plx_err=5e-5)
    args.update(kw)
    return make_problem(**args)


def make_problem_c4(seed=2298, **kw):
    """BASELINE config C4: 8192-pixel observed spectrum, 32 768-pixel model rows, 26 x 4 nodes.

    Model grid 5300 + 0.57*i A (i < 32768, to 23 977 A); the observed spectrum spans the whole
    model range at 4x coarser sampling; broadened at R = 1700 over
    that window; 2 contrast + 6 photometric synthetic passbands inside the model range.
    """
    rng = np.random.default_rng(seed + 11)
    n_data = int(kw.pop('n_data', 8192))
    step = (2.39765 - 0.53003) / (n_data - 1)
    data_wl = 0.53003 + step * np.arange(n_data)
    data_wl[1:-1] += rng.uniform(-0.2, 0.2, n_data - 2) * step          # slightly irregular sampling
    args = dict(specmin=5300.0, dlam=0.57, n_pix=32768, window_um=(0.5300, 2.3978), n_data=n_data,
                data_wl_um=data_wl, res=1700, seed=seed)
    args.update(kw)
    return make_problem(**args)


def _many_bands(wl, n_contrast, n_phot):
    span = wl[-1] - wl[0]
    cc = [synthetic_passband(wl[0] + (0.12 + 0.76 * k / max(n_contrast - 1, 1)) * span, 0.14 * span, 300 + 200 * k,
                             ripple_seed=30 + k) for k in range(n_contrast)]
    pc = [synthetic_passband(wl[0] + (0.06 + 0.88 * k / max(n_phot - 1, 1)) * span, (0.05 + 0.004 * k) * span,
                             200 + 60 * k, ripple_seed=50 + k) for k in range(n_phot)]
    cn = [f'synth_c{k}' for k in range(n_contrast)]
    pn = [('2MASS_' if k % 2 else 'SDSS_') + f'b{k}' for k in range(n_phot)]
    return cc, cn, pc, pn


def make_problem_3d_small(seed=2298, **kw):
    """Reduced 3-D (Teff, log g, [Fe/H]) problem for CPU-oracle parity: 5 x 4 x 3 nodes, 5000-pixel rows,
    4 contrast + 12 photometric bands."""
    wl = 5400.0 + 0.5 * np.arange(5000)
    cc, cn, pc, pn = _many_bands(wl, 4, 12)
    args = dict(temps=range(3500, 3901, 100), lgs=[4.0, 4.5, 5.0, 5.5], fehs=[-0.5, 0.0, 0.5], feh=0.2,
                specmin=5400.0, dlam=0.5, n_pix=5000, window_um=(0.60, 0.72), n_data=257, res=1700, seed=seed,
                contrast_curves=cc, contrast_names=cn, phot_curves=pc, phot_names=pn,
                truth=[3720.0, 3630.0, 0.10, 0.45, 0.7, 2.5e-3], plx=2.5e-3, plx_err=5e-5)
    args.update(kw)
    return make_problem(**args)


def make_problem_c5(seed=2298, **kw):
    """BASELINE config C5 (stress): 26 x 4 x 5 (Teff, log g, [Fe/H]) grid, 100 000-pixel rows (416 MB),
    25 000 observed pixels, 4 contrast + 12 photometric bands; the fit's metallicity is fixed at +0.1."""
    n_pix, dlam = 100000, 0.187
    wl = 5300.0 + dlam * np.arange(n_pix)
    cc, cn, pc, pn = _many_bands(wl, 4, 12)
    rng = np.random.default_rng(seed + 13)
    n_data = int(kw.pop('n_data', 25000))
    lo, hi = (wl[0] + 0.3) * 1e-4, (wl[-1] - 0.3) * 1e-4
    step = (hi - lo) / (n_data - 1)
    data_wl = lo + step * np.arange(n_data)
    data_wl[1:-1] += rng.uniform(-0.2, 0.2, n_data - 2) * step
    args = dict(fehs=[-1.0, -0.5, 0.0, 0.3, 0.5], feh=0.1, specmin=5300.0, dlam=dlam, n_pix=n_pix,
                window_um=(0.5300, (wl[-1] + 1) * 1e-4), n_data=n_data, data_wl_um=data_wl, res=1700, seed=seed,
                contrast_curves=cc, contrast_names=cn, phot_curves=pc, phot_names=pn)
    args.update(kw)
    return make_problem(**args)


def make_problem_triple_small(seed=2298, **kw):
    """Reduced triple-star problem (nspec = 3, 8 parameters): 6 x 4 nodes, 6000-pixel rows, 4 contrast
    filters (the first two constrain the secondary, the last two the tertiary: mft6.py:747-749)."""
    wl = 5400.0 + 0.5 * np.arange(6000)
    cc, cn, _, _ = _many_bands(wl, 4, 6)
    args = dict(nspec=3, contrast_curves=cc, contrast_names=cn, truth=[3800.0, 3650.0, 3480.0, 0.10, 0.45, 0.8, 0.6, 2.5e-3])
    args.update(kw)
    return make_problem_small(seed=seed, **args)


def sample_prior_box_triple(prob: Problem, n, seed=0):
    """theta batch for triples: [T1, T2, T3, Av, R1, R2/R1, R3/R1, plx] with a few forced edge cases."""
    rng = np.random.default_rng(seed)
    tmin, tmax = float(min(prob.temps)), float(max(prob.temps))
    th = np.empty((n, 8), dtype=np.float64)
    th[:, 0:3] = rng.uniform(tmin, tmax, (n, 3))
    th[:, 3] = rng.uniform(0.0, 1.0, n)
    th[:, 4] = rng.uniform(0.1, 1.8, n)                                   # no upper bound on R1 for triples
    th[:, 5:7] = rng.uniform(0.1, 1.0, (n, 2))
    th[:, 7] = rng.uniform(1 / 1000.0, 1 / 100.0, n)
    if n >= 8:
        th[0, 0] = tmin; th[1, 2] = tmax; th[2, 1] = float(prob.temps[1])
        th[3, 3] = 0.0; th[4, 3] = -0.01; th[5, 6] = 0.04; th[6, 7] = 1 / 1500.0; th[7, 1] = tmax + 1
    return th


def sample_prior_box(prob: Problem, n, seed=0, edge_cases=True):
    """theta batch uniform in the prior box of SURVEY.md section 8(d), with forced edge cases."""
    rng = np.random.default_rng(seed)
    tmin, tmax = float(min(prob.temps)), float(max(prob.temps))
    th = np.empty((n, 6), dtype=np.float64)
    th[:, 0] = rng.uniform(tmin, tmax, n)
    th[:, 1] = rng.uniform(tmin, tmax, n)
    th[:, 2] = rng.uniform(0.0, 1.0, n)
    th[:, 3] = rng.uniform(0.1, 1.2, n)
    th[:, 4] = rng.uniform(0.1, 1.0, n)
    th[:, 5] = rng.uniform(1 / 3000.0, 1 / 100.0, n)
    if edge_cases and n >= 16:
        nodes = prob.temps
        th[0, 0] = tmin; th[1, 0] = tmax; th[2, 1] = tmin; th[3, 1] = tmax
        th[4, 0] = float(nodes[len(nodes) // 2]); th[5, 1] = float(nodes[1])
        th[6, 0] = float(nodes[2]); th[6, 1] = float(nodes[2])          # both on the same node
        th[7, 2] = 0.0                                                    # Av = 0: extinction skipped
        th[8, 2] = -0.01                                                  # prior: -inf
        th[9, 0] = tmax + 1.0                                             # prior: -inf
        th[10, 3] = 0.04                                                  # prior: -inf
        th[11, 5] = 1 / 3.0                                               # prior: -inf
        th[12, 0] = 0.5 * (nodes[0] + nodes[1])                           # tie in find_nearest
        th[13, 5] = 1 / 3000.0                                            # far edge of parallax box
        th[14, 5] = 1 / 50.0                                              # nearer than first dust slice
        th[15, 3] = 1.5                                                   # R1 at its upper bound
    if getattr(prob, 'feh_free', False):
        # free-metallicity extension: trailing [Fe/H] column, uniform over the nodes, with node / edge / outside cases
        z = rng.uniform(min(prob.fehs), max(prob.fehs), n)
        if edge_cases and n >= 24:
            z[16] = prob.fehs[1]; z[17] = min(prob.fehs); z[18] = max(prob.fehs)
            z[19] = max(prob.fehs) + 0.01; z[20] = min(prob.fehs) - 0.01; z[21] = 0.5 * (prob.fehs[0] + prob.fehs[1])
        th = np.concatenate([th, z[:, None]], axis=1)
    return th
