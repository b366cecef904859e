"""CPU restatements of the third-party numerics that mft6.py calls on its hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import it, and only as the checker.  The product path (``mcmc_spec_b200``) never imports it.

PARITY UNPINNED.  None of the packages below is vendored under /root/reference, none is
version-pinned by the reference (no requirements file / lockfile), none is installed in the build
container, and the reference ships no golden vectors or known-answer tests for them.  Each
function restates the package's *published* algorithm and is pinned only by closed-form unit
tests (tests/test_oracle_closed_forms.py).

    package (unpinned)        reference call sites (mft6.py)        restated here as
    ------------------------  ------------------------------------  -------------------------
    extinction (kbarbary)     :62 ccm89, :63 apply                  ccm89(), apply()
    PyAstronomy.pyasl         :128 instrBroadGaussFast(maxsig=5)    instr_broad_gauss_fast()
    pyphot                    :21 get_library, :776 get_flux,       LibraryFilter.get_flux(),
                              :780/:782 Vega/AB_zero_flux           .lpivot / .AB_zero_flux
    dustmaps.bayestar         :23, :1233-1234 query(mode='samples') BayestarSightline.query()
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------------
# extinction.ccm89 / extinction.apply        (mft6.py:62-63)
# --------------------------------------------------------------------------------------------

def _ccm89_ab(x: np.ndarray):
    """a(x), b(x) of Cardelli, Clayton & Mathis (1989), x in inverse microns.

    Piecewise: IR power law (x < 1.1), optical/NIR 7th-order polynomial in y = x - 1.82
    (1.1 <= x < 3.3; CCM89 eq. 3a/3b), UV (3.3 <= x < 8) and far-UV (8 <= x <= 11).
    """
    x = np.asarray(x, dtype=np.float64)
    a = np.empty_like(x)
    b = np.empty_like(x)

    ir = x < 1.1
    y = x[ir] ** 1.61
    a[ir] = 0.574 * y
    b[ir] = -0.527 * y

    opt = (x >= 1.1) & (x < 3.3)
    y = x[opt] - 1.82
    a[opt] = ((((((0.32999 * y - 0.77530) * y + 0.01979) * y + 0.72085) * y - 0.02427) * y
               - 0.50447) * y + 0.17699) * y + 1.0
    b[opt] = ((((((-2.09002 * y + 5.30260) * y - 0.62251) * y - 5.38434) * y + 1.07233) * y
               + 2.28305) * y + 1.41338) * y

    uv = (x >= 3.3) & (x < 8.0)
    xu = x[uv]
    y = xu - 5.9
    fa = np.where(y > 0.0, -0.04473 * y * y - 0.009779 * y * y * y, 0.0)
    fb = np.where(y > 0.0, 0.2130 * y * y + 0.1207 * y * y * y, 0.0)
    a[uv] = 1.752 - 0.316 * xu - 0.104 / ((xu - 4.67) ** 2 + 0.341) + fa
    b[uv] = -3.090 + 1.825 * xu + 1.206 / ((xu - 4.62) ** 2 + 0.263) + fb

    fuv = x >= 8.0
    y = x[fuv] - 8.0
    a[fuv] = ((-0.070 * y + 0.137) * y - 0.628) * y - 1.073
    b[fuv] = ((0.374 * y - 0.420) * y + 4.257) * y + 13.670
    return a, b


def ccm89(wave, a_v, r_v=3.1):
    """A(lambda) in magnitudes for wavelengths in Angstrom: a_v * (a(x) + b(x) / r_v)."""
    wave = np.atleast_1d(np.asarray(wave, dtype=np.float64))
    x = 1.0e4 / wave
    if np.any(x < 0.3) or np.any(x > 11.0):
        raise ValueError("ccm89 is defined for 0.3 <= 1/lambda[um] <= 11")
    a, b = _ccm89_ab(x)
    return a_v * (a + b / r_v)


def apply(extinction_mag, flux):
    """extinction.apply: flux * 10**(-0.4 * A)."""
    trans = 10.0 ** (-0.4 * np.asarray(extinction_mag, dtype=np.float64))
    return np.asarray(flux, dtype=np.float64) * trans


# --------------------------------------------------------------------------------------------
# PyAstronomy.pyasl.instrBroadGaussFast(wvl, flux, resolution, maxsig=5)      (mft6.py:128)
# --------------------------------------------------------------------------------------------

def gauss_kernel_taps(dx: float, sigma: float, maxsig: float):
    """Sampled, sum-normalised Gaussian used by pyasl.broadGaussFast.

    ``lx = int(sigma*maxsig/dx*2) + 1`` taps at offsets ``(arange(lx) - sum(divmod(lx,2)) + 1)*dx``.
    """
    lx = int(((sigma * maxsig) / dx) * 2.0) + 1
    nx = (np.arange(lx, dtype=int) - sum(divmod(lx, 2)) + 1) * dx
    e = 1.0 / np.sqrt(2.0 * np.pi * sigma ** 2) * np.exp(-(0.0 - nx) ** 2 / (2.0 * sigma ** 2))
    e /= np.sum(e)
    return nx, e


def instr_broad_gauss_fast(wvl, flux, resolution, maxsig=5):
    wvl = np.asarray(wvl, dtype=np.float64)
    flux = np.asarray(flux, dtype=np.float64)
    dxs = wvl[1:] - wvl[:-1]
    if abs(dxs.max() - dxs.min()) > np.mean(dxs) * 1e-6:
        raise ValueError("The wavelength axis must be equidistant")
    fwhm = 1.0 / float(resolution) * np.mean(wvl)
    sigma = fwhm / (2.0 * np.sqrt(2.0 * np.log(2.0)))
    _, e = gauss_kernel_taps(dxs[0], sigma, maxsig)
    return np.convolve(flux, e, mode="same")


# --------------------------------------------------------------------------------------------
# pyphot library filter: get_flux / zero points                     (mft6.py:773-782)
# --------------------------------------------------------------------------------------------

C_AA_PER_S = 2.99792458e18  # speed of light in Angstrom / s


class LibraryFilter:
    """Stand-in for ``pyphot.get_library()[name]``.

    pyphot's own curves and Vega spectrum are in its HDF5 library, which is not in the reference
    repo; the curve is therefore an input (built from bps/ by the product's passband loader, or
    synthetic) and so is ``vega_zero_flux``.  ``dtype`` is 'photon' or 'energy' as in pyphot.
    """

    def __init__(self, name, wavelength, transmit, dtype="photon", vega_zero_flux=None):
        self.name = name
        self.wavelength = np.asarray(wavelength, dtype=np.float64)
        self.transmit = np.asarray(transmit, dtype=np.float64)
        self.dtype = dtype
        self.Vega_zero_flux = vega_zero_flux

    def reinterp(self, lamb):
        return np.interp(lamb, self.wavelength, self.transmit, left=0.0, right=0.0)

    def get_flux(self, slamb, sflux):
        slamb = np.asarray(slamb, dtype=np.float64)
        sflux = np.asarray(sflux, dtype=np.float64)
        ifT = self.reinterp(slamb)
        nonzero = np.where(ifT > 0)[0]
        if nonzero.size == 0:
            return 0.0
        lo = max(0, nonzero.min() - 5)
        hi = min(len(ifT), nonzero.max() + 5)
        lam = slamb[lo:hi]
        T = ifT[lo:hi]
        f = sflux[lo:hi]
        if "photon" in self.dtype:
            a = np.trapezoid(lam * T * f, lam)
            b = np.trapezoid(lam * T, lam)
        else:
            a = np.trapezoid(T * f, lam)
            b = np.trapezoid(T, lam)
        return a / b

    @property
    def lpivot(self):
        lam, T = self.wavelength, self.transmit
        if "photon" in self.dtype:
            return np.sqrt(np.trapezoid(lam * T, lam) / np.trapezoid(T / lam, lam))
        return np.sqrt(np.trapezoid(T, lam) / np.trapezoid(T / lam ** 2, lam))

    @property
    def AB_zero_mag(self):
        return 2.5 * np.log10(self.lpivot ** 2 / C_AA_PER_S) + 48.6

    @property
    def AB_zero_flux(self):
        return 10.0 ** (-0.4 * self.AB_zero_mag)


# --------------------------------------------------------------------------------------------
# dustmaps.bayestar.BayestarQuery(coords with distance, mode='samples')   (mft6.py:1233-1234)
# --------------------------------------------------------------------------------------------

class BayestarSightline:
    """One sky pixel of the Bayestar map: E(B-V)-like samples on a distance-modulus ladder.

    The reference queries a fixed (ra, dec) with the walker's distance, so a single sightline
    table ``samples[n_samples, n_dm]`` with bin edges ``dm_edges[n_dm]`` carries everything the
    prior needs.  Distance handling restates dustmaps' ``BayestarQuery.query``:
    nearer than the first slice -> first slice scaled by ``10**(0.2*(dm - dm0))``; beyond the last
    slice -> last slice; otherwise linear in distance modulus between the bracketing slices.
    """

    def __init__(self, dm_edges, samples):
        self.dm_edges = np.asarray(dm_edges, dtype=np.float64)
        self.samples = np.asarray(samples, dtype=np.float64)
        assert self.samples.shape[1] == self.dm_edges.shape[0]

    def query(self, distance_pc):
        d_kpc = distance_pc / 1.0e3
        dm = 5.0 * (np.log10(d_kpc) + 2.0)
        n = self.dm_edges.shape[0]
        ceil_idx = int(np.searchsorted(self.dm_edges, dm))
        if ceil_idx == 0:
            a = 10.0 ** (0.2 * (dm - self.dm_edges[0]))
            return a * self.samples[:, 0]
        if ceil_idx == n:
            return self.samples[:, -1].copy()
        hi, lo = self.dm_edges[ceil_idx], self.dm_edges[ceil_idx - 1]
        a = (hi - dm) / (hi - lo)
        return (1.0 - a) * self.samples[:, ceil_idx] + a * self.samples[:, ceil_idx - 1]
