"""Passband tables: the host-side mirror of ``get_transmission`` and of the pyphot library filters.

Reference: /root/reference/mft6.py:565-649 (``get_transmission``: filter name -> file under
``bps/`` -> (wavelength [A], throughput, n_resel, mean wavelength)), :3577-3597 (assembly of
``ctm`` / ``ptm`` and the global extrema ``tmi`` / ``tma``) and :766-783 (the pyphot library
filters that stand behind the photometric magnitudes).

The reference resolves names against files in ``bps/`` relative to the working directory; here
the directory is an argument.  All quirks of the name resolution are kept (``'Kp'`` is Keck K',
not Kepler; lp600 and Gaia in nm; keck_kp in microns; 2MASS Ks normalised to its maximum; FITS
columns stay float32 so that ``np.mean`` of them is a float32 mean, as astropy would deliver).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

C_AA_PER_S = 2.99792458e18
H_ERG_S = 6.62607015e-27


# ------------------------------------------------------------------------------ FITS BINTABLE

def read_fits_bintable(path, hdu=1):
    """Minimal FITS binary-table reader (scalar 'E', 'D', 'J', 'I' columns only).

    Returns {column name: array in the column's native precision}.  Stands in for
    ``astropy.io.fits.open(path)[hdu].data`` (mft6.py:616) -- every table the hot path needs
    (2MASS syn curves, Vega spectrum) is of this simple form.
    """
    with open(path, 'rb') as fh:
        raw = fh.read()
    off = 0
    for ihdu in range(hdu + 1):
        header = {}
        done = False
        while not done:
            block = raw[off:off + 2880]
            off += 2880
            if len(block) < 2880:
                raise ValueError(f"{path}: truncated FITS header")
            for i in range(0, 2880, 80):
                card = block[i:i + 80].decode('ascii', 'replace')
                if card.startswith('END'):
                    done = True
                    break
                if card[8:10] == '= ':
                    val = card[10:].split('/')[0].strip()
                    header[card[:8].strip()] = val.strip("'").strip()
        naxis = int(header.get('NAXIS', 0))
        nbytes = 0
        if naxis > 0:
            nbytes = abs(int(header['BITPIX'])) // 8
            for k in range(naxis):
                nbytes *= int(header[f'NAXIS{k + 1}'])
            nbytes += int(header.get('PCOUNT', 0))
        if ihdu < hdu:
            off += ((nbytes + 2879) // 2880) * 2880
    if header.get('XTENSION', '') != 'BINTABLE':
        raise ValueError(f"{path}: HDU {hdu} is not a BINTABLE")
    rowlen, nrows, nfields = int(header['NAXIS1']), int(header['NAXIS2']), int(header['TFIELDS'])
    codes = {'E': '>f4', 'D': '>f8', 'J': '>i4', 'I': '>i2'}
    fields = []
    for k in range(1, nfields + 1):
        form = header[f'TFORM{k}'].strip()
        rep = form[:-1]
        if rep not in ('', '1') or form[-1] not in codes:
            raise ValueError(f"{path}: unsupported TFORM {form!r}")
        fields.append((header[f'TTYPE{k}'].strip(), codes[form[-1]]))
    dt = np.dtype(fields)
    if dt.itemsize != rowlen:
        raise ValueError(f"{path}: row length mismatch")
    tab = np.frombuffer(raw, dtype=dt, count=nrows, offset=off)
    return {name: tab[name].astype(tab[name].dtype.newbyteorder('=')) for name, _ in fields}


# ------------------------------------------------------------------------------ get_transmission

def _split_name(f):
    """Filter-name parsing of mft6.py:570-594 (system guessed from the bare filter name)."""
    f = f.lower().strip(',')
    if ',' in f:
        syst, fil = f.split(',')
        return syst.strip(), fil.strip()
    fil = f
    syst = None
    if fil in 'i':
        syst = 'cousins'
    elif fil in 'ubvr':
        syst = 'johnson'
    elif fil == 'kp' or fil == 'kshort':
        syst = 'keck'
    elif fil in 'jhks':
        syst = '2mass'
    elif fil in '562 692 880':
        syst = 'dssi'
    elif fil in 'kepler':
        syst = 'kep'
    elif fil in 'brgamma':
        syst = 'nirc2'
    return syst, fil


def get_transmission(f, res, bps_dir):
    """Mirror of mft6.py:565-649.  Returns (t_wl [A], t_cv, n_resel, mean(t_wl))."""
    syst, fil = _split_name(f)
    P = lambda name: os.path.join(bps_dir, name)  # noqa: E731
    if fil == 'lp600':
        tab = np.genfromtxt(P('lp600.csv'), delimiter=',')
        t_wl, t_cv = tab[:, 0] * 10, tab[:, 1]
    elif syst == 'gaia' and fil in ('g', 'rp', 'bp'):
        t_wl, t_cv = np.genfromtxt(P(f'gaia_{fil}_pb.txt')).T
        t_wl = t_wl * 1e1
    elif syst == 'kep':
        t_wl, t_cv = np.genfromtxt(P('Kepler_Kepler.K.dat')).T
    elif syst == 'nirc2':
        t_wl, t_cv = np.genfromtxt(P('Keck_NIRC2.Brgamma.dat')).T
    elif syst == '2mass':
        if fil in ('j', 'h'):
            tab = read_fits_bintable(P(f'2mass_{fil}_001_syn.fits'))
            t_wl, t_cv = tab['WAVELENGTH'], tab['THROUGHPUT']
        elif fil in ('k', 'ks'):
            tab = np.genfromtxt(P('2MASS_2MASS.Ks.dat'))
            t_wl, t_cv = tab[:, 0], tab[:, 1] / max(tab[:, 1])
        else:
            raise KeyError(f"unknown 2MASS filter {fil!r}")
    elif syst == 'dssi':
        tab = np.genfromtxt(P(f'DSSI_{fil}nm.dat'))
        t_wl, t_cv = tab[:, 0], tab[:, 1]
    elif syst == 'sdss':
        t_wl, t_cv = np.genfromtxt(P(f'SLOAN_SDSS.{fil}prime_filter.dat')).T
    elif syst == 'keck' and fil == 'kp':
        tab = np.genfromtxt(P('keck_kp.txt'))
        t_wl, t_cv = tab[:, 0] * 1e4, tab[:, 1]
    elif syst == 'keck' and fil == 'kshort':
        t_wl, t_cv = np.genfromtxt(P('Keck_NIRC2.Ks.dat')).T
    else:
        tab = read_fits_bintable(P(f'{syst}_{fil}_002.fits'))
        t_wl, t_cv = tab['WAVELENGTH'], tab['THROUGHPUT']
    res_element = np.mean(t_wl) / res
    n_resel = (max(t_wl) - min(t_wl)) / res_element
    return t_wl, t_cv, n_resel, np.mean(t_wl)


def assemble_filter_tables(cfilt, pfilt, res, bps_dir):
    """ctm, ptm, tmi, tma exactly as mft6.py:3577-3597 builds them (lists of Python lists)."""
    tmi, tma = np.inf, 0

    def collect(names):
        nonlocal tmi, tma
        wls, tras, nre, cwl = [], [], [], []
        for f in names:
            w, t, re, c = get_transmission(f, res, bps_dir)
            wls.append(list(w)); tras.append(list(t)); nre.append(re); cwl.append(c)
            if min(w) < tmi:
                tmi = min(w)
            if max(w) > tma:
                tma = max(w)
        return [wls, tras, nre, cwl]

    ctm = collect(cfilt)
    ptm = collect(pfilt)
    return ctm, ptm, tmi, tma


# ------------------------------------------------------------------------------ library filters

@dataclass
class PhotBand:
    """One pyphot-library filter as the hot path needs it (mft6.py:773-782).

    ``wavelength`` [A] ascending, ``transmit``; ``dtype`` 'photon' or 'energy' (pyphot's detector
    type); ``system`` 'vega' -> magnitude against ``vega_zero_flux``, 'ab' -> ``ab_zero_flux``.
    """
    name: str
    wavelength: np.ndarray
    transmit: np.ndarray
    dtype: str = 'photon'
    system: str = 'vega'
    vega_zero_flux: float = float('nan')
    ab_zero_flux: float = float('nan')

    @property
    def zero_flux(self):
        return self.vega_zero_flux if self.system == 'vega' else self.ab_zero_flux


def band_flux(band: PhotBand, lam, flux):
    """pyphot ``UnitFilter.get_flux`` on (lam [A], flux): photon-type -> int(lam T f)/int(lam T)."""
    T = np.interp(lam, band.wavelength, band.transmit, left=0.0, right=0.0)
    if 'photon' in band.dtype:
        return np.trapezoid(lam * T * flux, lam) / np.trapezoid(lam * T, lam)
    return np.trapezoid(T * flux, lam) / np.trapezoid(T, lam)


def pivot_wavelength(band: PhotBand):
    lam, T = band.wavelength, band.transmit
    if 'photon' in band.dtype:
        return np.sqrt(np.trapezoid(lam * T, lam) / np.trapezoid(T / lam, lam))
    return np.sqrt(np.trapezoid(T, lam) / np.trapezoid(T / lam ** 2, lam))


def ab_zero_flux(band: PhotBand):
    """pyphot ``AB_zero_flux``: 10**(-0.4*(2.5 log10(lpivot^2 / c) + 48.6)) erg/s/cm^2/A."""
    return 10.0 ** (-0.4 * (2.5 * np.log10(pivot_wavelength(band) ** 2 / C_AA_PER_S) + 48.6))


def load_vega(path):
    """Data/vegaspec.fits (WAVELENGTH [A], FLUX [PHOTLAM]) -> (lam, f_lambda [erg/s/cm^2/A])."""
    tab = read_fits_bintable(path)
    lam = np.asarray(tab['WAVELENGTH'], dtype=np.float64)
    photlam = np.asarray(tab['FLUX'], dtype=np.float64)
    return lam, photlam * H_ERG_S * C_AA_PER_S / lam


# pyphot library names behind the photometry list (mft6.py:766-769) and the bps/ curve that
# substitutes for each (pyphot's own curves live in its HDF5 library, not in the reference repo).
LIBRARY_SUBSTITUTES = {
    'SDSS_r': ('sdss,r', 'ab'), 'SDSS_i': ('sdss,i', 'ab'), 'SDSS_z': ('sdss,z', 'ab'),
    '2MASS_J': ('j', 'vega'), '2MASS_H': ('h', 'vega'), '2MASS_Ks': ('k', 'vega'),
}


def reference_library_names(n_phot):
    """mft6.py:766-769."""
    if n_phot == 3:
        return ['2MASS_J', '2MASS_H', '2MASS_Ks']
    return ['SDSS_r', 'SDSS_i', 'SDSS_z', '2MASS_J', '2MASS_H', '2MASS_Ks']


def build_library_bands(n_phot, res, bps_dir, vega=None):
    """PhotBand list for the reference's photometric filters, zero points included.

    ``vega``: (lam, f_lambda) of the Vega spectrum; the Vega zero flux is the band flux of that
    spectrum (pyphot's definition).  Without it Vega-system zero points stay NaN.
    """
    bands = []
    for name in reference_library_names(n_phot)[:n_phot]:
        fname, system = LIBRARY_SUBSTITUTES[name]
        w, t, _, _ = get_transmission(fname, res, bps_dir)
        w = np.asarray(w, dtype=np.float64)
        t = np.asarray(t, dtype=np.float64)
        order = np.argsort(w, kind='mergesort')
        band = PhotBand(name, w[order], t[order], 'photon', system)
        band.ab_zero_flux = float(ab_zero_flux(band))
        if vega is not None:
            band.vega_zero_flux = float(band_flux(band, vega[0], vega[1]))
        bands.append(band)
    return bands
