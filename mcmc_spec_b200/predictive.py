"""Posterior-predictive sweep: Kepler / Gaia magnitudes of each component over posterior samples,
and the planet-radius correction factors derived from them.

Reference: ``plot_results`` draws 2000 rows of the flat chain without replacement and calls
``make_composite(..., plot=True)`` once per row (/root/reference/mft6.py:2486-2509); the plot
branch (:785-816) integrates each *component* spectrum through the Kepler curve
(``np.trapz(F_k * tran, wl)``, energy weighted, no zero point) and through pyphot's ``Gaia_G``
(``get_flux`` minus ``Vega_zero_mag``).  Then (:2515-2593)

    kep_contrast = sec_mag - pri_mag                    -> kep_contrast.txt
    gaia_sec, gaia_pri (component Gaia magnitudes)      -> gaia_sec.txt, gaia_pri.txt
    pri_corr = sqrt(1 + 10**(-0.4 kep_contrast))        -> pri_corr.txt   (Furlan+2017)
    sec_corr = ratio * sqrt(1 + 10**(0.4 kep_contrast)) -> sec_corr.txt

Here the 2000 forward models are ONE launch of the streaming kernel on a second handle that shares
the fit's device grid: the two bands are bound as per-component band sums (``build_tables(...,
component_bands=...)``) and read back through ``ms_component_mags``.  Binaries only (the reference's
``nspec == 3`` plot branch divides by a list, :820).  No CPU path: the handle needs the CUDA library.
"""
from __future__ import annotations

import os

import numpy as np

from . import native
from .passbands import PhotBand, band_flux, get_transmission
from .tables import build_tables


def kepler_curve(res, bps_dir):
    """``get_transmission('kepler', res)`` (mft6.py:788): (ran [A], tm)."""
    w, t, _, _ = get_transmission('kepler', res, bps_dir)
    return np.asarray(w, dtype=np.float64), np.asarray(t, dtype=np.float64)


def gaia_g_band(res, bps_dir, vega=None):
    """Stand-in for pyphot ``lib['Gaia_G']`` (mft6.py:811) from ``bps/gaia_g_pb.txt``; photon type,
    Vega zero flux = band flux of the Vega spectrum when ``vega`` = (lam, f_lambda) is given."""
    w, t, _, _ = get_transmission('gaia,g', res, bps_dir)
    w, t = np.asarray(w, dtype=np.float64), np.asarray(t, dtype=np.float64)
    o = np.argsort(w, kind='mergesort')
    band = PhotBand('Gaia_G', w[o], t[o], 'photon', 'vega')
    if vega is not None:
        band.vega_zero_flux = float(band_flux(band, vega[0], vega[1]))
    return band


class ComponentSweep:
    """Per-component Kepler and Gaia magnitudes for batches of samples, on the GPU."""

    def __init__(self, problem, kepler, gaia: PhotBand, device=0, share_grid_with=None):
        if problem.nspec != 2:
            raise ValueError("the posterior-predictive sweep is defined for binaries (mft6.py:796-816)")
        self.problem, self.gaia = problem, gaia
        self.tables = build_tables(problem, component_bands=[('energy', kepler[0], kepler[1]), ('library', gaia)])
        owner = None if share_grid_with is None else share_grid_with.engine
        self.engine = native.Engine(self.tables, device=device, grid_owner=owner)

    def __call__(self, theta):
        """theta [n, 6] -> dict of [n] arrays: pri_mag, sec_mag (Kepler, instrumental), gaia_pri, gaia_sec,
        gaiamag (composite)."""
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        mags = self.engine.component_mags_host(theta)                     # [n, 2 bands, 2 stars]
        zero_mag = -2.5 * np.log10(self.gaia.vega_zero_flux)              # pyphot Vega_zero_mag
        gp, gs = mags[:, 1, 0] - zero_mag, mags[:, 1, 1] - zero_mag       # mft6.py:813-814
        comp = 10.0 ** (-0.4 * mags[:, 1, 0]) + 10.0 ** (-0.4 * mags[:, 1, 1])   # get_flux is linear in the spectrum
        return dict(pri_mag=mags[:, 0, 0], sec_mag=mags[:, 0, 1], gaia_pri=gp, gaia_sec=gs,
                    gaiamag=-2.5 * np.log10(comp / self.gaia.vega_zero_flux))      # mft6.py:812

    def close(self):
        self.engine.close()


def correction_factors(kep_contrast, kep_rad):
    """mft6.py:2531-2532."""
    kep_contrast, kep_rad = np.asarray(kep_contrast), np.asarray(kep_rad)
    return np.sqrt(1 + 10 ** (-0.4 * kep_contrast)), kep_rad * np.sqrt(1 + 10 ** (0.4 * kep_contrast))


def percentile_summary(x):
    """|50th|, |84th|, |16th| percentile as the reference titles its histograms (mft6.py:2517-2519)."""
    return tuple(float(np.abs(np.percentile(np.asarray(x), q))) for q in (50, 84, 16))


def kepler_sweep(sweep: ComponentSweep, sample, size=2000, rng=None):
    """mft6.py:2486-2512 for the 6-parameter ``dist_fit=True`` chain: draw ``size`` rows of the flat chain
    without replacement (``rng``: numpy Generator or RandomState; the reference uses the global
    ``np.random``), forward-model them in one launch, derive contrasts and correction factors."""
    sample = np.asarray(sample, dtype=np.float64)
    if sample.ndim != 2 or sample.shape[1] != 6:
        raise ValueError("sample must be the flat chain [n, 6]")
    rng = np.random if rng is None else rng
    rows = sample[rng.choice(len(sample), size=int(min(size, len(sample))), replace=False), :]
    m = sweep(rows)
    kep_contrast = m['sec_mag'] - m['pri_mag']
    kep_rad = rows[:, 4]                                                   # ratio1 = rad2 (:2492)
    pri_corr, sec_corr = correction_factors(kep_contrast, kep_rad)
    return dict(sample=rows, kep_contrast=kep_contrast, kep_rad=kep_rad, gaia_sec=m['gaia_sec'],
                gaia_pri=m['gaia_pri'], gaiamag=m['gaiamag'], pri_corr=pri_corr, sec_corr=sec_corr)


def write_sweep(run_dir, res):
    """The five text files ``plot_results`` leaves in the run directory (mft6.py:2527-2529, :2563, :2582)."""
    os.makedirs(run_dir, exist_ok=True)
    for name, key in (('kep_contrast.txt', 'kep_contrast'), ('gaia_sec.txt', 'gaia_sec'), ('gaia_pri.txt', 'gaia_pri'),
                      ('pri_corr.txt', 'pri_corr'), ('sec_corr.txt', 'sec_corr')):
        np.savetxt(os.path.join(run_dir, name), np.array(res[key]))
