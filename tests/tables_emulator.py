"""numpy walk-through of the device algorithm on the host tables (TEST infrastructure).

Mirrors walker_prep_kernel + lnprob_stream_kernel step for step so the table builder and the
kernel's arithmetic plan can be checked against the oracle on a machine without a GPU.  It is not
shipped and not importable from the package.
"""
import numpy as np

RSUN, PC, SIGMA_SB, LSUN = 6.957e10, 3.086e18, 5.670374e-5, 3.839e33


def _interp(x, y, v):
    if not (x[0] <= v <= x[-1]):
        return None
    return float(np.interp(v, x, y))


def _bracket(nodes, v):
    i1 = int(np.abs(nodes - v).argmin())
    if nodes[i1] == v:
        return i1, i1, 0.0
    i2 = i1 - 1 if nodes[i1] > v else i1 + 1
    if i2 < 0 or i2 >= len(nodes):
        return None
    return i1, i2, (v - nodes[i1]) / (nodes[i2] - nodes[i1])


def _dense(packed, n_reg, tile):
    rows = []
    for f in range(packed.n):
        w = np.zeros(n_reg)
        lo, hi = int(packed.tile_lo[f]) * tile, int(packed.tile_hi[f]) * tile
        seg = packed.weights[int(packed.offset[f]):int(packed.offset[f]) + hi - lo]
        w[lo:min(hi, n_reg)] = seg[:min(hi, n_reg) - lo]
        rows.append(w)
    return rows


def lnprob(t, theta, tile=1024):
    ns = t.nspec
    p = np.asarray(theta, dtype=np.float64)
    if not np.all(np.isfinite(p)):
        return -np.inf
    temps, av, rad, plx = p[:ns], p[ns], p[ns + 1:2 * ns + 1], p[2 * ns + 1]
    # logprior's own `rad`: p0[nspec:2*nspec] in the dist_fit branch when ext is False (mft6.py:1219-1221)
    prad = p[ns:2 * ns] if (ns == 2 and t.dist_fit and not t.ext_flag) else rad
    out = any(temps > t.tmax) or any(temps < t.tmin)
    if t.dist_fit:
        out = out or any(prad < 0.05) or prad[0] > 1.5 or plx < 1 / 3000 or plx > 1 / 4
    else:
        out = out or rad[1] < 0.05 or rad[0] < 0.05
    if t.ext_flag and av < 0:
        out = True
    if out:
        return -np.inf
    logg, lum = [], []
    for T in temps:
        g = _interp(t.iso_teff, t.iso_logg, T); l = _interp(t.iso_teff, t.iso_lum, T)
        if g is None:
            return -np.inf
        logg.append(g); lum.append(l)
    lp = 0.0
    if t.ext_flag:
        dm = 5.0 * (np.log10((1 / plx) / 1e3) + 2.0)
        c = int(np.searchsorted(t.dust_dm_edges, dm))
        n = len(t.dust_dm_edges)
        if c == 0:
            s = 10.0 ** (0.2 * (dm - t.dust_dm_edges[0])) * t.dust_samples[:, 0]
        elif c == n:
            s = t.dust_samples[:, -1]
        else:
            a = (t.dust_dm_edges[c] - dm) / (t.dust_dm_edges[c] - t.dust_dm_edges[c - 1])
            s = (1 - a) * t.dust_samples[:, c] + a * t.dust_samples[:, c - 1]
        s = s * 3.1 * 0.884
        mu, sig = s.mean(), s.std()
        sig = 0.05 if sig == 0 else sig
        lp += -0.5 * ((av - mu) / sig) ** 2
    for k in range(len(p) if t.dist_fit else len(p) - 2):
        if t.prior_mu[k] != 0:
            lp += -0.5 * ((p[k] - t.prior_mu[k]) / t.prior_sig[k]) ** 2
    if t.rad_prior:
        rm = [np.sqrt(lum[s] * LSUN / (4 * np.pi * SIGMA_SB * temps[s] ** 4)) / RSUN for s in range(2)]
        for val, m in ((prad[0], rm[0]), (prad[1], rm[1] / rm[0])):
            lp += -0.5 * ((val - m) / (0.02 * m)) ** 2
    grid = t.grid.reshape(-1, t.row_len)[:, t.reg_lo:t.reg_lo + t.n_reg]
    comps = []
    di = 1.0 / plx
    for s in range(ns):
        bt = _bracket(t.teff_nodes, temps[s]); bg = _bracket(t.logg_nodes, logg[s])
        if bt is None or bg is None:
            return -np.inf
        base = rad[0] * RSUN / (di * PC) if s == 0 else rad[0] * rad[s] * RSUN / (di * PC)
        scale = base * base
        f = np.zeros(t.n_reg)
        for it, ct in ((bt[0], 1 - bt[2]), (bt[1], bt[2])):
            for ig, cg in ((bg[0], 1 - bg[2]), (bg[1], bg[2])):
                f += (ct * cg * scale) * grid[ig * len(t.teff_nodes) + it]
        comps.append(f)
    comp = comps[0] + comps[1]
    av_eff = av if (t.ext_flag and av > 0) else 0.0
    cw = _dense(t.contrast, t.n_reg, tile); pw = _dense(t.phot, t.n_reg, tile)
    hi = t.rebin_hi; lo = hi - 1
    ylo = comp[lo] * np.exp2(av_eff * t.ext_lo); yhi = comp[hi] * np.exp2(av_eff * t.ext_hi)
    m = ylo + (yhi - ylo) * t.rebin_t                      # unscaled model on the observed pixels
    B = t.poly_basis
    x = t.data_x
    q = np.stack([np.full_like(x, B[0, 0]), B[1, 0] + B[1, 1] * x, B[2, 0] + x * (B[2, 1] + B[2, 2] * x)])
    a = q @ (t.data_flux / m)                               # fit to data/model, unscaled
    chi_un = np.sum((m - t.data_flux / (a @ q)) ** 2 * (1.0 / (t.data_err * t.data_err)))
    s = t.data_median / np.median(m)                        # the median scaling factors out: chi2 = s^2 * chi_un
    iic = s * s * chi_un / len(m)
    icon = 0.0
    for f in range(t.contrast.n):
        con = -2.5 * np.log10(cw[f] @ comps[1]) - (-2.5 * np.log10(cw[f] @ comps[0]))
        icon += (con - t.c_mag[f]) ** 2 / t.c_err[f] ** 2
    iph = 0.0
    for f in range(t.phot.n):
        mag = -2.5 * np.log10((pw[f] @ comp) * t.p_norm[f]) + av_eff * t.p_ext[f]
        iph += (mag - t.p_mag[f]) ** 2 / t.p_err[f] ** 2
    cs = iic * (t.contrast.n + t.phot.n) + icon + iph
    return lp + (-np.inf if np.isnan(cs) else -0.5 * cs)
