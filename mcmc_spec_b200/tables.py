"""Walker-independent tables: Problem (reference shapes) -> flat arrays for the device.

Everything that does not depend on the walker is folded into tables once, on the host, in
float64, so that the kernels only stream grid rows and do dot products:

  * the consumed region of the model grid          -- make_composite's get_spec cut, mft6.py:687, :542
  * per contrast filter: throughput interpolated onto the model pixels (interp1d semantics, stable
    sort of the curve) times the trapezoid quadrature weights of the in-band pixels
                                                     -- mft6.py:717-731
  * per photometric (library) filter: lam*T (photon) or T (energy) times quadrature weights, the
    normalising integral, the zero point, and the CCM89 coefficient at the mean wavelength
                                                     -- mft6.py:771-783, :1163; pyphot get_flux
  * per observed pixel: the bracketing model pixels, the interpolation fraction, and the CCM89
    exponent coefficients of both brackets          -- mft6.py:1162, :1169-1170
  * an orthonormal quadratic basis on the mapped abscissae for norm_spec's least squares
                                                     -- mft6.py:193-196, numpy Polynomial.fit
  * the isochrone columns as interp1d would sort them -- mft6.py:70-76, :90-95
  * the prior vector in (mu, sigma) per parameter    -- mft6.py:1241-1260

TILE is the pixel tile of the streaming kernel; filter weights are stored zero-padded on whole
tiles so the kernel needs no bounds checks.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

TILE = 1024
LOG2_10 = np.log2(10.0)


def ccm89_k(wl_aa, rv=3.1):
    """a(x) + b(x)/Rv of CCM89 for wavelengths in Angstrom (A(lambda) = Av * k).  x <= 3.3 only
    is needed by this path (model grids start at 3000 A) but the UV branches are kept."""
    x = 1.0e4 / np.atleast_1d(np.asarray(wl_aa, dtype=np.float64))
    if np.any(x < 0.3) or np.any(x > 11.0):
        raise ValueError("wavelength outside the CCM89 validity range")
    a = np.empty_like(x); b = np.empty_like(x)
    m = x < 1.1
    p = x[m] ** 1.61
    a[m], b[m] = 0.574 * p, -0.527 * p
    m = (x >= 1.1) & (x < 3.3)
    y = x[m] - 1.82
    a[m] = ((((((0.32999 * y - 0.77530) * y + 0.01979) * y + 0.72085) * y - 0.02427) * y - 0.50447) * y
            + 0.17699) * y + 1.0
    b[m] = ((((((-2.09002 * y + 5.30260) * y - 0.62251) * y - 5.38434) * y + 1.07233) * y + 2.28305) * y
            + 1.41338) * y
    m = (x >= 3.3) & (x < 8.0)
    xu = x[m]; y = xu - 5.9
    fa = np.where(y > 0, -0.04473 * y ** 2 - 0.009779 * y ** 3, 0.0)
    fb = np.where(y > 0, 0.2130 * y ** 2 + 0.1207 * y ** 3, 0.0)
    a[m] = 1.752 - 0.316 * xu - 0.104 / ((xu - 4.67) ** 2 + 0.341) + fa
    b[m] = -3.090 + 1.825 * xu + 1.206 / ((xu - 4.62) ** 2 + 0.263) + fb
    m = x >= 8.0
    y = x[m] - 8.0
    a[m] = ((-0.070 * y + 0.137) * y - 0.628) * y - 1.073
    b[m] = ((0.374 * y - 0.420) * y + 4.257) * y + 13.670
    return a + b / rv


def trapz_weights(x):
    """w with sum(w*y) == np.trapezoid(y, x)."""
    x = np.asarray(x, dtype=np.float64)
    w = np.zeros_like(x)
    if x.size >= 2:
        d = np.diff(x)
        w[:-1] += 0.5 * d
        w[1:] += 0.5 * d
    return w


def _sorted_curve(w, t):
    """interp1d(assume_sorted=False): stable (mergesort) argsort of the abscissae."""
    w = np.asarray(w, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    o = np.argsort(w, kind='mergesort')
    return w[o], t[o]


@dataclass
class PackedFilters:
    n: int
    tile_lo: np.ndarray          # [n] first tile carrying weights
    tile_hi: np.ndarray          # [n] one past the last tile
    offset: np.ndarray           # [n] element offset of the filter's weights in ``weights``
    weights: np.ndarray          # concatenated, whole tiles, zero padded
    support: list = field(default_factory=list)   # (first, last+1) region pixels with weight


def _pack(weight_rows, n_reg):
    """Dense region-length weight rows -> tile-aligned packed storage."""
    tl, th, off, chunks, sup = [], [], [], [], []
    pos = 0
    for w in weight_rows:
        nz = np.nonzero(w)[0]
        if nz.size == 0:
            tl.append(0); th.append(0); off.append(pos); sup.append((0, 0))
            continue
        lo_t, hi_t = nz[0] // TILE, nz[-1] // TILE + 1
        buf = np.zeros((hi_t - lo_t) * TILE, dtype=np.float64)
        a, b = lo_t * TILE, min(hi_t * TILE, n_reg)
        buf[:b - a] = w[a:b]
        tl.append(lo_t); th.append(hi_t); off.append(pos); chunks.append(buf); sup.append((nz[0], nz[-1] + 1))
        pos += buf.size
    weights = np.concatenate(chunks) if chunks else np.zeros(1)
    return PackedFilters(len(weight_rows), np.array(tl, np.int32), np.array(th, np.int32),
                         np.array(off, np.int64), weights, sup)


@dataclass
class DeviceTables:
    # grid / region
    teff_nodes: np.ndarray
    logg_nodes: np.ndarray
    feh_nodes: np.ndarray
    feh: float
    grid: np.ndarray             # the Problem's grid (host, C-contiguous), rows of length P
    row_len: int                 # P
    reg_lo: int                  # first consumed pixel
    n_reg: int                   # consumed pixels per row
    wl_reg: np.ndarray           # wavelengths of the consumed pixels (host-side bookkeeping)
    # observed spectrum
    n_data: int
    rebin_hi: np.ndarray         # int32 [N]
    rebin_t: np.ndarray          # [N]
    ext_lo: np.ndarray           # [N]  E = exp2(Av * ext)
    ext_hi: np.ndarray           # [N]
    data_flux: np.ndarray
    data_err: np.ndarray
    data_x: np.ndarray           # abscissae mapped to [-1, 1]
    poly_basis: np.ndarray       # [3,3] q_k(x) = sum_j B[k,j] x^j, orthonormal over data_x
    data_median: float
    # filters
    contrast: PackedFilters
    c_mag: np.ndarray
    c_err: np.ndarray
    phot: PackedFilters
    p_norm: np.ndarray           # 1 / (int(lam T) * zero_flux)
    p_ext: np.ndarray            # CCM89 k at the filter's mean wavelength
    p_mag: np.ndarray
    p_err: np.ndarray
    # isochrone / dust / prior
    iso_teff: np.ndarray
    iso_logg: np.ndarray
    iso_lum: np.ndarray
    dust_dm_edges: np.ndarray
    dust_samples: np.ndarray     # [n_samples, n_dm]
    tmin: float
    tmax: float
    prior_mu: np.ndarray         # [ndim]  (0 -> no Gaussian prior on that parameter)
    prior_sig: np.ndarray
    nspec: int
    ext_flag: int
    dist_fit: int
    rad_prior: int
    flavour: int = 0             # 1: mft6_nospec.py
    feh_free: int = 0            # 1: [Fe/H] is the trailing walker parameter (3-D grids)

    @property
    def n_tiles(self):
        return (self.n_reg + TILE - 1) // TILE

    def bytes_per_eval(self):
        """Algorithmic bytes of one evaluation:
        n_star * 2^D * P_reg * 8 + 8*ndim + 8."""
        ndim = (6 if self.nspec == 2 else 8) + int(self.feh_free)
        corners = 4 if self.feh_nodes.size <= 1 else 8
        return self.nspec * corners * self.n_reg * 8 + 8 * ndim + 8


def orthonormal_quadratic_basis(x):
    """B[3,3] (lower triangular) with q_k(x) = sum_j B[k,j] x^j and sum_i q_k(x_i) q_l(x_i) = delta_kl.

    Gram-Schmidt in extended precision; the least-squares quadratic through (x_i, y_i) is then
    p(x) = sum_k (sum_i q_k(x_i) y_i) q_k(x): the same polynomial Polynomial.fit returns
    (mft6.py:195), without an N x 3 SVD per walker.
    """
    xl = np.asarray(x, dtype=np.longdouble)
    V = np.stack([np.ones_like(xl), xl, xl * xl])          # monomials
    B = np.zeros((3, 3), dtype=np.longdouble)
    Q = []
    for k in range(3):
        c = np.zeros(3, dtype=np.longdouble); c[k] = 1
        v = V[k].copy()
        for _ in range(2):                                   # re-orthogonalise once
            for l, q in enumerate(Q):
                pr = np.sum(q * v)
                v = v - pr * q
                c = c - pr * B[l]
        nrm = np.sqrt(np.sum(v * v))
        Q.append(v / nrm)
        B[k] = c / nrm
    return np.asarray(B, dtype=np.float64)


def build_tables(prob, component_bands=None) -> DeviceTables:
    """``component_bands``: optional list replacing the contrast filters by bands whose per-component
    flux is wanted (the posterior-predictive sweep, mft6.py:785-816): ``('energy', wl, tm)`` -- the
    contrast-type integral ``trapz(F_k * tran, wl)`` of mft6.py:798 -- or ``('library', PhotBand)`` --
    pyphot ``get_flux(F_k)`` (:813).  The consumed region stays the one of the fit's own filters."""
    wl = np.asarray(prob.wl, dtype=np.float64)
    data_wl, data_flux = (np.asarray(a, dtype=np.float64) for a in prob.data)
    err = np.asarray(prob.err, dtype=np.float64)
    wls, tras, _, _ = prob.ctm
    phot_wls, _, _, phot_cwl = prob.ptm

    # ---- consumed region (mft6.py:663-673, :687, :539-542)
    wlmin, wlmax = np.inf, 0
    for w in list(wls) + list(phot_wls):
        wlmin = min(wlmin, min(w)); wlmax = max(wlmax, max(w))
    r = [min(data_wl), max(data_wl)]
    region = [min(min(r), prob.tmi / 1e4, wlmin / 1e4) - 1e-4, max(max(r), prob.tma / 1e4, wlmax / 1e4) + 1e-4]
    reg = np.array(region) * 1e4
    keep = np.where((wl >= min(reg)) & (wl <= max(reg)))[0]
    if keep.size < 2 or not np.all(np.diff(keep) == 1):
        raise ValueError("consumed region is empty or not contiguous")
    reg_lo, n_reg = int(keep[0]), int(keep.size)
    wl_reg = wl[keep]

    # ---- contrast filters (mft6.py:717-731)
    crows = []
    for w, t in zip(wls, tras):
        ws, ts = _sorted_curve(w, t)
        inb = np.where((wl_reg <= ws[-1]) & (wl_reg >= ws[0]))[0]
        row = np.zeros(n_reg)
        if inb.size:
            row[inb] = np.interp(wl_reg[inb], ws, ts) * trapz_weights(wl_reg[inb])
        crows.append(row)
    c_mag, c_err = np.array(prob.fr[0], dtype=np.float64), np.array(prob.fr[1], dtype=np.float64)
    if component_bands is not None:
        crows = []
        for spec in component_bands:
            row = np.zeros(n_reg)
            if spec[0] == 'energy':
                ws, ts = _sorted_curve(spec[1], spec[2])
                inb = np.where((wl_reg <= ws[-1]) & (wl_reg >= ws[0]))[0]
                if inb.size:
                    row[inb] = np.interp(wl_reg[inb], ws, ts) * trapz_weights(wl_reg[inb])
            elif spec[0] == 'library':
                band = spec[1]
                T = np.interp(wl_reg, band.wavelength, band.transmit, left=0.0, right=0.0)
                nz = np.nonzero(T > 0)[0]
                if nz.size:
                    lo, hi = max(0, nz.min() - 5), min(n_reg, nz.max() + 5)
                    lam = wl_reg[lo:hi]
                    kern = lam * T[lo:hi] if 'photon' in band.dtype else T[lo:hi]
                    row[lo:hi] = kern * trapz_weights(lam) / np.trapezoid(kern, lam)
            else:
                raise ValueError("component band kind must be 'energy' or 'library'")
            crows.append(row)
        # no observed contrast belongs to these bands: infinite errors keep them out of the chi-square
        c_mag, c_err = np.zeros(len(crows)), np.full(len(crows), np.inf)
    contrast = _pack(crows, n_reg)

    # ---- photometric library filters (pyphot get_flux; mft6.py:771-783)
    prows, p_norm, p_ext = [], [], []
    for band, cw in zip(prob.photbands, phot_cwl):
        T = np.interp(wl_reg, band.wavelength, band.transmit, left=0.0, right=0.0)
        nz = np.nonzero(T > 0)[0]
        row = np.zeros(n_reg)
        if nz.size:
            lo, hi = max(0, nz.min() - 5), min(n_reg, nz.max() + 5)
            lam = wl_reg[lo:hi]
            kern = lam * T[lo:hi] if 'photon' in band.dtype else T[lo:hi]
            row[lo:hi] = kern * trapz_weights(lam)
            denom = np.trapezoid(kern, lam)
        else:
            denom = np.nan
        prows.append(row)
        p_norm.append(1.0 / (denom * band.zero_flux))
        p_ext.append(float(ccm89_k(float(cw))[0]))
    phot = _pack(prows, n_reg)

    # ---- rebin + extinction tables (mft6.py:1162, :1169-1170; np.interp bracket semantics)
    x_new = data_wl * 1e4
    if x_new.min() < wl_reg[0] or x_new.max() > wl_reg[-1]:
        raise ValueError("A value in x_new is outside the interpolation range.")   # interp1d
    hi = np.searchsorted(wl_reg, x_new, side='right')
    hi = np.clip(hi, 1, n_reg - 1).astype(np.int32)
    lo = hi - 1
    rebin_t = (x_new - wl_reg[lo]) / (wl_reg[hi] - wl_reg[lo])
    kreg = ccm89_k(wl_reg)
    ext_coef = -0.4 * LOG2_10 * kreg
    order_ok = np.all(np.diff(hi) >= 0)
    if not order_ok:
        raise ValueError("observed wavelengths must be ascending")

    # ---- norm_spec abscissae (numpy Polynomial.fit domain mapping, polyutils.mapparms)
    xmin, xmax = data_wl.min(), data_wl.max()
    scl = 2.0 / (xmax - xmin)
    off = (xmax * -1.0 - xmin * 1.0) / (xmax - xmin)
    data_x = off + scl * data_wl
    basis = orthonormal_quadratic_basis(data_x)

    # ---- isochrone columns (mft6.py:70-76, :90-95) in interp1d's sorted order
    m = prob.matrix
    sel = np.where(np.array(m[:, 1]) == 9.0)
    teff5, logg5, lum5 = m[:, 4][sel][:220], m[:, 5][sel][:220], m[:, 6][sel][:220]
    o = np.argsort(teff5, kind='mergesort')

    # ---- prior vector (mft6.py:1241-1260 / :1298-1315)
    ns = prob.nspec
    ndim = 2 * ns + 2                       # the reference's parameters (a free [Fe/H] column carries no Gaussian prior)
    mu = np.zeros(ndim); sig = np.ones(ndim)
    pr = prob.prior
    if not (np.isscalar(pr) and pr == 0):
        pr = [float(v) for v in pr]
        if prob.dist_fit:
            ps = pr[:ns] + [pr[2 * ns]] + pr[2 * ns + 2:3 * ns + 2] + [pr[-2]]
            ss = pr[ns:2 * ns] + [pr[2 * ns + 1]] + pr[3 * ns + 2:4 * ns + 2] + [pr[-1]]
        else:
            ps = pr[:ns] + [pr[2 * ns]] + pr[2 * ns + 2:3 * ns + 1]
            ss = pr[ns:2 * ns] + [pr[2 * ns + 1]] + pr[3 * ns + 2:4 * ns + 1]
        for k, p in enumerate(ps):
            if p != 0:
                mu[k], sig[k] = p, ss[k]

    fehs = np.array(prob.fehs if prob.fehs is not None else [0.0], dtype=np.float64)
    grid = np.ascontiguousarray(prob.grid, dtype=np.float64)
    return DeviceTables(
        teff_nodes=np.array(prob.temps, dtype=np.float64), logg_nodes=np.array(prob.lgs, dtype=np.float64),
        feh_nodes=fehs, feh=float(prob.feh), grid=grid, row_len=int(wl.size), reg_lo=reg_lo, n_reg=n_reg,
        wl_reg=wl_reg, n_data=int(data_wl.size), rebin_hi=hi, rebin_t=rebin_t,
        ext_lo=ext_coef[lo], ext_hi=ext_coef[hi], data_flux=data_flux, data_err=err, data_x=data_x,
        poly_basis=basis, data_median=float(np.median(data_flux)),
        contrast=contrast, c_mag=c_mag, c_err=c_err,
        phot=phot, p_norm=np.array(p_norm), p_ext=np.array(p_ext),
        p_mag=np.array(prob.fr[3], dtype=np.float64), p_err=np.array(prob.fr[4], dtype=np.float64),
        iso_teff=np.ascontiguousarray(teff5[o]), iso_logg=np.ascontiguousarray(logg5[o]),
        iso_lum=np.ascontiguousarray(lum5[o]), dust_dm_edges=np.asarray(prob.dust_dm_edges, dtype=np.float64),
        dust_samples=np.ascontiguousarray(prob.dust_samples, dtype=np.float64),
        tmin=float(min(prob.temps)), tmax=float(max(prob.temps)), prior_mu=mu, prior_sig=sig, nspec=ns,
        ext_flag=int(bool(prob.av)), dist_fit=int(bool(prob.dist_fit)), rad_prior=int(bool(prob.rad_prior)),
        flavour=int(bool(getattr(prob, 'nospec', False))), feh_free=int(bool(getattr(prob, 'feh_free', False))))
