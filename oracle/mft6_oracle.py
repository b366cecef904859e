"""numpy/scipy restatement of the per-walker log-posterior of mft6.py (the parity oracle).

TEST INFRASTRUCTURE ONLY -- see oracle/thirdparty.py for the import rule.  The reference script
cannot be imported offline (emcee, PyAstronomy, extinction, pyphot, synphot, astropy, dustmaps,
corner, matplotlib and the BT-Settl / MIST / bayestar2019 data are all absent), and it ships no
golden vectors for this path, so this module *is* the statement of truth for the parity tests.

PARITY UNPINNED for the third-party numerics (oracle/thirdparty.py) in the sense that the reference
holds no vectors for them and cannot be run; they are held instead to published known answers from
outside this repository (tests/test_published_known_answers.py: the extinction package's documented
ccm89 vector, CCM89 Table 3, PyAstronomy's FWHM / sigma / tap rule, the 3631 Jy AB zero point, the
Cohen et al. 2003 2MASS zero points, dustmaps' distance rule).  numpy/scipy calls (``interp1d``,
``Polynomial.fit``, ``np.median``, ``np.trapezoid``) are the real libraries.

Every function names the reference lines it follows (``/root/reference/mft6.py``).  Argument
order follows the reference where that is meaningful; arguments the reference passes but never
reads (``ndust, broadening, vs, w, dust, norm, mode``) are dropped.  Two things the reference
gets from the environment are explicit inputs here:

  * ``temps`` / ``lgs``: the sorted node lists that ``get_spec`` rebuilds from ``glob`` on the
    model directory at every call (mft6.py:423-436, :457-463);
  * ``photlib``: the pyphot filter library (mft6.py:21) and ``dustmap``: the Bayestar query
    object (mft6.py:23).
"""
from __future__ import annotations

import numpy as np
from scipy.interpolate import interp1d

from . import thirdparty as tp

SIGMA_SB = 5.670374e-5   # erg/s/cm^2/K^4   (mft6.py:79)
LSUN = 3.839e33          # erg/s            (mft6.py:80)
RSUN = 6.957e10          # cm               (mft6.py:81, :691)
PC_CM = 3.086e18         # cm               (mft6.py:691)


# ---------------------------------------------------------------------------- small helpers

def extinct(wl, spec, av, rv=3.1):
    """mft6.py:46-64 -- CCM89 curve for (av, rv) turned into a transmission and applied."""
    ext_mag = tp.ccm89(wl, av, rv)
    return np.array(tp.apply(ext_mag, spec))


def _isochrone_columns(matrix, col):
    """Rows of the MIST table with log-age exactly 9.0, first 220 (mft6.py:70-76, :90-95)."""
    aage = matrix[:, 1]
    sel = np.where(np.array(aage) == 9.0)
    return matrix[:, 4][sel][:220], matrix[:, col][sel][:220]


def get_radius(teff, matrix):
    """mft6.py:66-85 -- L(Teff) from the 1 Gyr isochrone, then Stefan-Boltzmann radius [Rsun]."""
    teff5, lum5 = _isochrone_columns(matrix, 6)
    lum = interp1d(teff5, lum5)(teff)
    return np.sqrt(lum * LSUN / (4 * np.pi * SIGMA_SB * teff ** 4)) / RSUN


def get_logg(teff, matrix):
    """mft6.py:87-98 -- log g(Teff) from the same isochrone."""
    teff5, logg5 = _isochrone_columns(matrix, 5)
    return interp1d(teff5, logg5)(teff)


def find_nearest(array, value):
    """mft6.py:100-113 -- argmin |array - value| (ties -> lowest index)."""
    return int(np.abs(np.asarray(array) - value).argmin())


def chisq(model, data, var):
    """mft6.py:115-122 -- (model - data)^2 / var^2, elementwise (``var`` is a 1-sigma error)."""
    return ((np.array(model) - np.array(data)) ** 2) / np.array(var) ** 2


def broaden(even_wl, modelspec_interp, res):
    """mft6.py:124-152 (vsini = limb = 0) -- instrumental Gaussian + the 5/10-sample edge patch."""
    broad = tp.instr_broad_gauss_fast(even_wl, modelspec_interp, res, maxsig=5)
    broad[0:5] = broad[5]
    broad[len(broad) - 10:len(broad)] = broad[len(broad) - 11]
    return np.array(even_wl), np.array(broad)


def norm_spec(wl, model, data):
    """mft6.py:193-196 -- divide the data by a quadratic fitted to data/model."""
    frac = data / model
    p_fitted = np.polynomial.Polynomial.fit(wl, frac, deg=2)
    return data / p_fitted(wl)


def interp_2_spec(spec1, spec2, ep1, ep2, val):
    """mft6.py:198-209 -- pixelwise straight line through (ep1, spec1), (ep2, spec2) at val."""
    return ((np.array(spec2) - np.array(spec1)) / (ep2 - ep1)) * (val - ep1) + np.array(spec1)


# ---------------------------------------------------------------------------- grid assembly

def splice_broadened_grid(native, wl, window, resolution):
    """mft6.py:366-383 applied to rows already on the common 0.2 A grid.

    ``native``: dict '{T}, {g}' -> row on ``wl``.  For each row the data window
    ``min(window) <= wl <= max(window)`` is broadened (``broaden``) and spliced between the
    untouched blue and red parts.  Returns the ``specs`` dict including the ``'wl'`` entry.
    (The text-file reading and the interp1d resampling of mft6.py:343-371 sit upstream of this.)
    """
    wl = np.asarray(wl)
    inwin = np.where((wl >= min(window)) & (wl <= max(window)))
    specs = {}
    ww = None
    for k, spflux in native.items():
        ww, brd = broaden(wl[inwin], spflux[inwin], resolution)
        specs[k] = np.concatenate((spflux[np.where(wl < min(window))], brd,
                                   spflux[np.where(wl > max(window))]))
    specs['wl'] = np.array(np.concatenate((wl[np.where(wl < min(window))], ww,
                                           wl[np.where(wl > max(window))])))
    return specs


def _bracket(nodes, value):
    """Nearest node and its partner on the far side of ``value`` (mft6.py:439-453, :467-477)."""
    i1 = find_nearest(nodes, value)
    if nodes[i1] == value:
        i2 = i1
    elif nodes[i1] > value:
        i2 = i1 - 1            # NB: -1 wraps to the last node when value is below the grid
    else:
        i2 = i1 + 1            # IndexError when value is above the grid, as in the reference
    return nodes[i1], nodes[i2]


def get_spec(temp, log_g, reg, specdict, temps, lgs):
    """mft6.py:387-563, btsettl branch, wlunit='aa', without the per-call glob.

    Returns (wavelength [A], flux) cut to ``min(reg) <= wl <= max(reg)`` (``reg`` in microns).
    """
    temp1, temp2 = _bracket(temps, temp)
    lg1, lg2 = _bracket(lgs, log_g)
    spwave = specdict['wl']

    key = '{}, {}'.format
    if lg1 == lg2 and temp1 == temp2:
        spflux = specdict[key(temp1, lg1)]
    else:
        spec1 = specdict[key(temp1, lg1)]
        spec2 = specdict[key(temp2, lg2)]
        t1_inter = specdict[key(temp1, lg2)]
        t2_inter = specdict[key(temp2, lg1)]
        if lg1 != lg2 and temp1 != temp2:
            t1_lg = interp_2_spec(spec1, t1_inter, lg1, lg2, log_g)
            t2_lg = interp_2_spec(t2_inter, spec2, lg1, lg2, log_g)
            spflux = interp_2_spec(t1_lg, t2_lg, temp1, temp2, temp)
        elif lg1 == lg2 and temp1 != temp2:
            spflux = interp_2_spec(spec1, spec2, temp1, temp2, temp)
        else:
            spflux = interp_2_spec(spec1, spec2, lg1, lg2, log_g)

    reg = np.array(reg) * 1e4
    keep = np.where((spwave >= min(reg)) & (spwave <= max(reg)))
    return np.array(spwave[keep]), np.array(spflux[keep])


# ---------------------------------------------------------------------------- forward model

def reference_photlib_names(n_phot):
    """mft6.py:766-769 -- which pyphot library filters stand behind the photometry list."""
    if n_phot == 3:
        return ['2MASS_J', '2MASS_H', '2MASS_Ks']
    return ['SDSS_r', 'SDSS_i', 'SDSS_z', '2MASS_J', '2MASS_H', '2MASS_Ks']


def make_composite(teff, logg, rad, distance, contrast_filt, phot_filt, r, specs, ctm, ptm,
                   tmi, tma, temps, lgs, photlib, nspec=2, fs=None):
    """mft6.py:651-831, plot=False.

    Returns (wl [A], composite flux, contrast list, photometric central wavelengths, magnitudes).
    ``photlib`` maps pyphot library names to thirdparty.LibraryFilter; ``fs`` overrides the
    reference's fixed name list (only needed for >6 photometric bands, an extension).
    """
    wls, tras, n_res_el, cwl = ctm
    phot_wls, phot_tras, phot_resel, phot_cwl = ptm

    wlmin, wlmax = np.inf, 0
    for w in list(wls) + list(phot_wls):
        if min(w) < wlmin:
            wlmin = min(w)
        if max(w) > wlmax:
            wlmax = max(w)

    region = [min(min(r), tmi / 1e4, wlmin / 1e4) - 1e-4, max(max(r), tma / 1e4, wlmax / 1e4) + 1e-4]

    pri_wl, pri_spec = get_spec(teff[0], logg[0], region, specs, temps, lgs)
    fit_distance = not type(distance) == bool
    if fit_distance:
        di = 1 / distance
        pri_spec *= (rad[0] * RSUN / (di * PC_CM)) ** 2

    wl_rows, flux_rows = [pri_wl], [pri_spec]
    for n in range(1, len(teff)):
        sec_wl, sec_spec = get_spec(teff[n], logg[n], region, specs, temps, lgs)
        if fit_distance:
            di = 1 / distance
            sec_spec = sec_spec * (rad[0] * rad[n] * RSUN / (di * PC_CM)) ** 2
        else:
            sec_spec = sec_spec * rad[n - 1] ** 2
        wl_rows.append(sec_wl)
        flux_rows.append(sec_spec)
    pri_wl = np.vstack(wl_rows)
    pri_spec = np.vstack(flux_rows)

    # contrasts: energy-weighted band flux of each *component*, unextincted   (:717-741)
    mag = np.zeros((len(contrast_filt), len(teff)))
    for n in range(len(contrast_filt)):
        ran, tm = wls[n], tras[n]
        inband = np.where((pri_wl[0, :] <= max(ran)) & (pri_wl[0, :] >= min(ran)))
        w = pri_wl[0, :][inband]
        tran = interp1d(ran, tm)(w)
        for k in range(len(teff)):
            s = pri_spec[k][inband]
            m = np.trapezoid(s * tran, w)
            mag[n][k] = -2.5 * np.log10(m)

    if float(nspec) == 2:
        contrast = [mag[n][1] - mag[n][0] for n in range(len(contrast_filt))]
        spec1 = pri_spec[0, :] + pri_spec[1, :]
    else:
        half = int(len(contrast_filt) / 2)
        contrast1 = [mag[n][1] - mag[n][0] for n in range(len(contrast_filt))]
        contrast2 = [mag[n][2] - mag[n][0] for n in range(len(contrast_filt))]
        contrast = list(np.concatenate((contrast1[:half], contrast2[half:])))
        spec1 = pri_spec[0, :] + pri_spec[1, :] + pri_spec[2, :]

    # unresolved photometry through the library filters                         (:766-783)
    if fs is None:
        fs = reference_photlib_names(len(phot_filt))
    phot_phot = []
    for n in range(len(phot_filt)):
        f = photlib[fs[n]]
        flux = f.get_flux(pri_wl[0, :], spec1)
        if '2MASS' in fs[n]:
            mags = -2.5 * np.log10(flux / f.Vega_zero_flux)
        else:
            mags = -2.5 * np.log10(flux / f.AB_zero_flux)
        phot_phot.append(mags)

    return (np.array(pri_wl[0, :]), np.array(spec1), [c for c in contrast],
            np.array([float(p) for p in phot_cwl]), np.array([p for p in phot_phot]))


def make_composite_plot(teff, logg, rad, distance, r, specs, ctm, ptm, tmi, tma, temps, lgs,
                        kepler_curve, gaia_lib):
    """mft6.py:651-716 + :785-816 -- ``make_composite(..., plot=True)`` for binaries: the component
    spectra on the consumed region, each component's energy-weighted Kepler band magnitude
    (``np.trapz(F_k * tran, wl)``, no zero point) and its pyphot ``Gaia_G`` Vega magnitude
    ``-2.5 log10(get_flux(F_k)) - Vega_zero_mag``; plus the composite's Gaia magnitude.

    ``kepler_curve``: (ran, tm) of ``get_transmission('kepler', res)``; ``gaia_lib``: the
    ``lib['Gaia_G']`` stand-in (thirdparty.LibraryFilter with ``Vega_zero_flux``).  The reference
    also interpolates the ``'gaia,g'`` bps curve (:789, :794) but never uses the result.  The
    ``nspec == 3`` branch divides by an undefined-at-that-point list ``zp`` (:820) and is not restated.

    Returns (wl, spec1, pri_spec, sec_spec, pri_mag, sec_mag, gaia_pri, gaia_sec, gaiamag).
    """
    wls, _, _, _ = ctm
    phot_wls, _, _, _ = ptm
    wlmin, wlmax = np.inf, 0
    for w in list(wls) + list(phot_wls):
        wlmin = min(wlmin, min(w)); wlmax = max(wlmax, max(w))
    region = [min(min(r), tmi / 1e4, wlmin / 1e4) - 1e-4, max(max(r), tma / 1e4, wlmax / 1e4) + 1e-4]

    pri_wl, pri_spec = get_spec(teff[0], logg[0], region, specs, temps, lgs)
    sec_wl, sec_spec = get_spec(teff[1], logg[1], region, specs, temps, lgs)
    if not type(distance) == bool:
        di = 1 / distance
        pri_spec = pri_spec * (rad[0] * RSUN / (di * PC_CM)) ** 2
        sec_spec = sec_spec * (rad[0] * rad[1] * RSUN / (di * PC_CM)) ** 2
    else:
        sec_spec = sec_spec * rad[0] ** 2                                   # mft6.py:703
    spec1 = pri_spec + sec_spec

    ran, tm = kepler_curve
    inband = np.where((pri_wl >= min(ran)) & (pri_wl <= max(ran)))
    data_tm = interp1d(ran, tm)(pri_wl[inband])
    pri_phot = np.trapezoid(pri_spec[inband] * data_tm, pri_wl[inband])    # :798
    sec_phot = np.trapezoid(sec_spec[inband] * data_tm, pri_wl[inband])    # :799
    pri_mag, sec_mag = -2.5 * np.log10(pri_phot), -2.5 * np.log10(sec_phot)

    vega_zero_mag = -2.5 * np.log10(gaia_lib.Vega_zero_flux)                # pyphot Vega_zero_mag
    gaiamag = -2.5 * np.log10(gaia_lib.get_flux(pri_wl, spec1) / gaia_lib.Vega_zero_flux)   # :812
    gaia_pri = -2.5 * np.log10(gaia_lib.get_flux(pri_wl, pri_spec)) - vega_zero_mag         # :813
    gaia_sec = -2.5 * np.log10(gaia_lib.get_flux(pri_wl, sec_spec)) - vega_zero_mag         # :814
    return pri_wl, spec1, pri_spec, sec_spec, pri_mag, sec_mag, gaia_pri, gaia_sec, gaiamag


# ---------------------------------------------------------------------------- posterior

def loglikelihood(p0, fr, nspec, data, err, r, specs, ctm, ptm, tmi, tma, matrix, temps, lgs,
                  photlib, av=True, optimize=False, fs=None, parts=None, nospec=False):
    """mft6.py:1139-1205 for len(p0) in (6, 8).  ``nospec=True`` gives mft6_nospec.py:1140-1206, where
    every line of the spectral term is commented out and the total is contrast + photometry only.

    ``parts``: optional dict that receives the intermediate quantities (for parity debugging).
    """
    wl, spec = np.array(data)
    if len(p0) not in (6, 8):
        raise ValueError("p0 must have 6 or 8 entries")
    t_guess, extinct_guess = p0[:nspec], p0[nspec]
    rad_guess, dist_guess = p0[nspec + 1:2 * nspec + 1], p0[2 * nspec + 1]
    lg_guess = [get_logg(t, matrix) for t in t_guess]
    wave1, init_cspec, contrast, phot_cwl, phot = make_composite(
        t_guess, lg_guess, rad_guess, dist_guess, fr[2], fr[5], r, specs, ctm, ptm, tmi, tma,
        temps, lgs, photlib, nspec=nspec, fs=fs)

    if av == True and extinct_guess > 0:          # noqa: E712  (as in the reference)
        if not nospec:
            init_cspec = extinct(wave1, init_cspec, extinct_guess)
        init_phot = -2.5 * np.log10(extinct(phot_cwl, 10 ** (-0.4 * phot), extinct_guess))
    else:
        init_phot = phot

    chi_contrast = chisq(contrast, fr[0], fr[1])
    icontrast = np.sum(chi_contrast)
    chi_phot = chisq(init_phot, fr[3], fr[4])
    iphot = np.sum(chi_phot)
    if nospec:
        iic = np.nan
        init_cs = np.sum((icontrast, iphot))                           # mft6_nospec.py:1192
    else:
        init_cspec = interp1d(wave1, init_cspec)(wl * 1e4)
        init_cspec *= np.median(spec) / np.median(init_cspec)
        spec = norm_spec(wl, init_cspec, spec)
        ic = chisq(init_cspec, spec, err)
        iic = np.sum(ic) / len(ic)
        init_cs = np.sum((iic * (len(chi_contrast) + len(chi_phot)), icontrast, iphot))

    if parts is not None:
        parts.update(model=init_cspec, data_renorm=spec, chi2_spec=iic, contrast=np.array(contrast),
                     phot=np.array(init_phot), chi2_contrast=icontrast, chi2_phot=iphot,
                     logg=np.array(lg_guess, dtype=float))
    if optimize:
        return init_cs
    if np.isnan(init_cs):
        return -np.inf
    return -0.5 * init_cs


def logprior(p0, nspec, tmin, tmax, matrix, dustmap, prior=0, ext=True, dist_fit=True,
             rad_prior=False, plx_max=1 / 4):
    """mft6.py:1207-1327 -- the two 6-parameter branches (binary; dist_fit True / False).

    ``dustmap.query(distance_pc)`` stands for ``bayestar(SkyCoord(ra, dec, distance), 'samples')``.
    ``plx_max``: 1/4 in mft6.py:1227, 1/100 in mft6_nospec.py:1228.
    """
    temps = p0[:nspec]
    if len(p0) == 8:
        return _logprior_triple(p0, nspec, tmin, tmax, matrix, dustmap, prior, ext, dist_fit, rad_prior)
    if len(p0) != 6:
        raise ValueError("P0 doesn't match what I was expecting")          # mft6.py:1456-1457 prints this
    pp = []

    def av_term(extinct_val, plx):
        av_guess = dustmap.query(1 / plx) * 3.1 * 0.884
        mu, sig = np.mean(av_guess), np.std(av_guess)
        if sig == 0:
            sig = 0.05
        return -0.5 * ((extinct_val - mu) / sig) ** 2

    if dist_fit == True:                           # noqa: E712
        if ext == True:                            # noqa: E712
            extinct_val = p0[nspec]
            rad = p0[nspec + 1:2 * nspec + 1]
        else:
            rad = p0[nspec:2 * nspec]
        dist = p0[2 * nspec + 1]
        if (any(t > tmax for t in temps) or any(t < tmin for t in temps)
                or any(r < 0.05 for r in rad) or rad[0] > 1.5 or dist < 1 / 3000 or dist > plx_max):
            return -np.inf
        if ext == True and extinct_val < 0:        # noqa: E712
            return -np.inf
        elif ext == True:                          # noqa: E712
            pp.append(av_term(extinct_val, dist))

        if not (np.isscalar(prior) and prior == 0):
            prior = list(prior)
            ps = prior[:nspec] + [prior[2 * nspec]] + prior[2 * nspec + 2:3 * nspec + 2] + [prior[-2]]
            ss = (prior[nspec:2 * nspec] + [prior[2 * nspec + 1]] + prior[3 * nspec + 2:4 * nspec + 2]
                  + [prior[-1]])
            for k, p in enumerate(ps):
                if p != 0:
                    pp.append(-0.5 * ((p0[k] - p) / ss[k]) ** 2)
        rad_for_prior = rad
    else:
        rad, rad1 = p0[-2], p0[-3]
        if ext == True:                            # noqa: E712
            extinct_val = p0[-4]
        if any(t > tmax for t in temps) or any(t < tmin for t in temps) or rad < 0.05 or rad1 < 0.05:
            return -np.inf
        if ext == True and extinct_val < 0:        # noqa: E712
            return -np.inf
        elif ext == True:                          # noqa: E712
            pp.append(av_term(extinct_val, p0[-1]))

        if not (np.isscalar(prior) and prior == 0):
            prior = list(prior)
            ps = prior[:nspec] + [prior[2 * nspec]] + prior[2 * nspec + 2:3 * nspec + 1]
            ss = prior[nspec:2 * nspec] + [prior[2 * nspec + 1]] + prior[3 * nspec + 2:4 * nspec + 1]
            for k, p in enumerate(ps):
                if p != 0:
                    pp.append(-0.5 * ((p0[k] - p) / ss[k]) ** 2)
        rad_for_prior = [rad1, rad]

    if rad_prior == True:                          # noqa: E712
        model_radius1 = get_radius(temps[0], matrix)
        model_radius2 = get_radius(temps[1], matrix)
        for k, p in enumerate([model_radius1, model_radius2 / model_radius1]):
            pp.append(-0.5 * ((rad_for_prior[k] - p) / (0.02 * p)) ** 2)
    return np.sum(pp)


def _logprior_triple(p0, nspec, tmin, tmax, matrix, dustmap, prior, ext, dist_fit, rad_prior):
    """mft6.py:1329-1454 -- the two 8-parameter (triple) branches.  Quirks kept: no upper bound on R1 and a
    parallax floor of 1/1000 with a distance (:1346); without one only the two ratios and the sign of the
    parallax are checked (:1408).  With ``dist_fit`` the reference returns its sum only inside
    ``if rad_prior`` (:1381-1393); ``rad_prior=False`` there falls off the end of the function (None ->
    TypeError in logposterior), which is reported here as a ValueError."""
    temps = p0[:nspec]
    rad1, rad2, rad, dist = p0[-3], p0[-2], p0[-4], p0[-1]
    extinct_val = p0[nspec] if dist_fit else p0[-5]
    pp = []
    if dist_fit == True:                           # noqa: E712
        if (any(t > tmax for t in temps) or any(t < tmin for t in temps)
                or any(r < 0.05 for r in [rad, rad1, rad2]) or dist < 1 / 1000 or dist > 1 / 4):
            return -np.inf
    else:
        if any(t > tmax for t in temps) or any(t < tmin for t in temps) or rad1 < 0.05 or rad2 < 0.05 or dist < 0:
            return -np.inf
    if ext == True and extinct_val < 0:            # noqa: E712
        return -np.inf
    elif ext == True:                              # noqa: E712
        mu, sig = _dust_moments(dustmap, dist)
        pp.append(-0.5 * ((extinct_val - mu) / sig) ** 2)
    if not (np.isscalar(prior) and prior == 0):
        prior = list(prior)
        if dist_fit == True:                       # noqa: E712
            ps = prior[:nspec] + [prior[2 * nspec]] + prior[2 * nspec + 2:3 * nspec + 2] + [prior[-2]]
            ss = (prior[nspec:2 * nspec] + [prior[2 * nspec + 1]] + prior[3 * nspec + 2:4 * nspec + 2]
                  + [prior[-1]])
        else:
            ps = prior[:nspec] + [prior[2 * nspec]] + prior[2 * nspec + 2:3 * nspec + 1]
            ss = prior[nspec:2 * nspec] + [prior[2 * nspec + 1]] + prior[3 * nspec + 2:4 * nspec + 1]
        for k, p in enumerate(ps):
            if p != 0:
                pp.append(-0.5 * ((p0[k] - p) / ss[k]) ** 2)
    if rad_prior == True:                          # noqa: E712
        m1, m2, m3 = (get_radius(temps[k], matrix) for k in range(3))
        for val, p in zip([rad, rad1, rad2], [m1, m2 / m1, m3 / m1]):
            pp.append(-0.5 * ((val - p) / (0.02 * p)) ** 2)
    elif dist_fit == True:                         # noqa: E712
        raise ValueError("mft6.py:1329-1393 returns None for a triple with dist_fit and without rad_prior")
    return np.sum(pp)


def logposterior(p0, fr, nspec, data, err, r, specs, ctm, ptm, tmi, tma, tmin, tmax, matrix,
                 temps, lgs, photlib, dustmap, prior=0, a=True, dist_fit=True, rad_prior=False,
                 fs=None, parts=None, nospec=False):
    """mft6.py:1459-1470 -- prior first; a non-finite prior short-circuits to -inf.
    ``nospec=True``: the same function of mft6_nospec.py."""
    lp = logprior(p0, nspec, tmin, tmax, matrix, dustmap, prior=prior, ext=a, dist_fit=dist_fit,
                  rad_prior=rad_prior, plx_max=(1 / 100 if nospec else 1 / 4))
    if parts is not None:
        parts['lp'] = lp
    if not np.isfinite(lp):
        return -np.inf
    lh = loglikelihood(p0, fr, nspec, data, err, r, specs, ctm, ptm, tmi, tma, matrix, temps,
                       lgs, photlib, av=a, fs=fs, parts=parts, nospec=nospec)
    return lp + lh


# ---------------------------------------------------------------------------- optimiser stage

def opt_prior(vals, pval, psig):
    """mft6.py:833-854 -- sum of ((value - prior)/sigma)^2 over the priors that are defined."""
    pp = []
    if len(pval) == 1 or type(pval) == float:
        try:
            pp.append(((float(vals) - float(pval[0])) / float(psig[0])) ** 2)
        except TypeError:
            pp.append(((vals[0] - pval[0]) / psig[0]) ** 2)
    else:
        for k, p in enumerate(pval):
            if p != 0:
                pp.append(((vals[k] - pval[k]) / (psig[k])) ** 2)
    return np.sum(pp)


def _dust_moments(dustmap, plx):
    av_guess = dustmap.query(1 / plx) * 3.1 * 0.884
    mu, sig = np.mean(av_guess), np.std(av_guess)
    return mu, (0.05 if sig == 0 else sig)


def fit_spec_init(wl_aa, flux, err, reg, t_guess, av, rad_guess, fr_guess, specs, distance, ctm, ptm, tmi, tma,
                  matrix, temps, lgs, photlib, dustmap, nspec=2, dist_fit=True, rad_prior=False, fs=None, nospec=False):
    """mft6.py:859-930 (distance given): the starting chi-square of one optimiser walker and the observed
    spectrum after its one-off renormalisation.  ``wl_aa`` is the data wavelength in Angstrom (the
    reference multiplies its argument by 1e4 in place, :860).  Returns (init_cs, flux_renormalised).

    Quirks kept: the initial model is *not* extincted (:880 commented out) and the initial photometric
    chi-square uses the unextincted magnitudes (:902 uses ``phot``, not ``init_phot``)."""
    dist, pprior, psig = distance
    extinct_guess, eprior, esig = av
    lg_guess = [get_logg(t, matrix) for t in t_guess]
    wave1, init_cspec, contrast, phot_cwl, phot = make_composite(
        t_guess, lg_guess, rad_guess, dist, fr_guess[2], fr_guess[5], reg, specs, ctm, ptm, tmi, tma,
        temps, lgs, photlib, nspec=nspec, fs=fs)
    chi_contrast = chisq(contrast, fr_guess[0], fr_guess[1])
    ip = chisq(phot, fr_guess[3], fr_guess[4])
    if nospec:                                     # mft6_nospec.py:885-905: the spectral lines are commented out
        init_cs = np.sum((np.sum(chi_contrast), np.sum(ip)))
    else:
        init_cspec = interp1d(wave1, init_cspec)(wl_aa)
        init_cspec *= np.median(flux) / np.median(init_cspec)
        flux = norm_spec(wl_aa, init_cspec, flux)
        iic = np.sum(chisq(init_cspec, flux, err)) / len(flux) * 3
        init_cs = np.sum((iic * (len(chi_contrast) + len(ip)), np.sum(chi_contrast), np.sum(ip)))
    if dist_fit == True:                           # noqa: E712
        init_cs += opt_prior([dist], [pprior], [psig])
    if rad_prior == True:                          # noqa: E712
        r1, r2 = get_radius(t_guess[0], matrix), get_radius(t_guess[1], matrix)
        if nspec == 2:                             # mft6.py:917-918
            init_cs += opt_prior(rad_guess, [r1, r2 / r1], [0.05 * r for r in rad_guess])
        elif nspec == 3:                           # mft6.py:920-922
            r3 = get_radius(t_guess[2], matrix)
            init_cs += opt_prior(rad_guess, [r1, r2 / r1, r3 / r1], [0.05 * r for r in rad_guess])
    mu, sig = _dust_moments(dustmap, dist)
    init_cs += opt_prior([extinct_guess], [mu], [sig])
    return init_cs, flux


def fit_spec_test_cs(var_par, wl_aa, flux, err, reg, fr_guess, specs, pprior, psig, rad_sigma, ctm, ptm, tmi, tma,
                     matrix, temps, lgs, photlib, dustmap, nspec=2, dist_fit=True, rad_prior=False, fs=None, nospec=False):
    """mft6.py:989-1046 -- chi-square of one in-box proposal ``var_par = [T[2], Av, rad[2], plx]`` against
    the renormalised ``flux``; ``rad_sigma`` is ``si[2]`` (:1038)."""
    mu, sig = _dust_moments(dustmap, var_par[3])
    logg_guess = [get_logg(v, matrix) for v in var_par[0]]
    test_wave1, test_cspec, test_contrast, test_phot_cwl, test_phot = make_composite(
        var_par[0], logg_guess, var_par[2], float(var_par[3]), fr_guess[2], fr_guess[5], reg, specs, ctm, ptm,
        tmi, tma, temps, lgs, photlib, nspec=nspec, fs=fs)
    if var_par[1] > 0:
        if not nospec:
            test_cspec = extinct(test_wave1, test_cspec, var_par[1])
        test_phot = -2.5 * np.log10(extinct(test_phot_cwl, 10 ** (-0.4 * test_phot), var_par[1]))
    chi_contrast = chisq(test_contrast, fr_guess[0], fr_guess[1])
    chi_phot = chisq(test_phot, fr_guess[3], fr_guess[4])
    if nospec:                                     # mft6_nospec.py:1005-1030
        test_cs = np.sum((np.sum(chi_contrast), np.sum(chi_phot)))
    else:
        test_cspec = interp1d(test_wave1, test_cspec)(wl_aa)
        test_cspec *= np.median(flux) / np.median(test_cspec)
        tc = chisq(test_cspec, flux, err)
        ttc = np.sum(tc) / len(tc) * 3
        test_cs = np.sum((ttc * (len(chi_contrast) + len(chi_phot)), np.sum(chi_contrast), np.sum(chi_phot)))
    test_cs += opt_prior(var_par[1], [mu], [sig])
    if dist_fit == True:                           # noqa: E712
        test_cs += opt_prior([var_par[3]], [pprior], [psig])
    if rad_prior == True and nspec == 2:           # noqa: E712
        r1, r2 = get_radius(var_par[0][0], matrix), get_radius(var_par[0][1], matrix)
        test_cs += opt_prior(var_par[2], [r1, r2 / r1], np.array(rad_sigma))
    elif rad_prior == True and nspec == 3:         # noqa: E712   mft6.py:1044-1050
        r1, r2, r3 = (get_radius(var_par[0][k], matrix) for k in range(3))
        test_cs += opt_prior(var_par[2], [r1, r2 / r1, r3 / r1], np.array(rad_sigma))
    return test_cs


# ---------------------------------------------------------------------------- problem-level glue

class OracleProblem:
    """Binds the reference-shaped inputs of one fit so tests can call ``lnprob(theta)``.

    ``prob`` is any object with the attributes the reference's ``main`` assembles
    (mft6.py:3492-3597, :3689): ``fr, nspec, data, err, specs, ctm, ptm, tmi, tma, matrix,
    temps, lgs, photlib, dustmap, prior, av, dist_fit, rad_prior`` and optionally ``fs``.
    """

    def __init__(self, prob):
        self.p = prob
        data_wl = prob.data[0]
        self.r = [min(data_wl), max(data_wl)]                       # mft6.py:3688
        self.tmin, self.tmax = min(prob.temps), max(prob.temps)      # mft6.py:1479-1487

    def lnprob(self, p0, parts=None):
        p = self.p
        return logposterior(np.asarray(p0, dtype=np.float64), p.fr, p.nspec, p.data, p.err, self.r,
                            p.specs, p.ctm, p.ptm, p.tmi, p.tma, self.tmin, self.tmax, p.matrix,
                            p.temps, p.lgs, p.photlib, p.dustmap, prior=p.prior, a=p.av,
                            dist_fit=p.dist_fit, rad_prior=p.rad_prior,
                            fs=getattr(p, 'fs', None), parts=parts, nospec=getattr(p, 'nospec', False))

    def lnprob_batch(self, theta):
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        return np.array([self.lnprob(t) for t in theta], dtype=np.float64)
