"""Read/write helpers for tests/golden/*.npz (see tests/golden/make_fixtures.py)."""
import numpy as np

from mcmc_spec_b200 import synthetic
from mcmc_spec_b200.passbands import PhotBand


def pack_curves(prefix, wls, tras):
    """Ragged (wl, T) curves -> flat arrays + lengths, dtype preserved per curve via a flag."""
    lens = np.array([len(w) for w in wls], dtype=np.int64)
    is32 = np.array([np.asarray(w).dtype == np.float32 for w in wls])
    wl = np.concatenate([np.asarray(w, dtype=np.float64) for w in wls])
    tr = np.concatenate([np.asarray(t, dtype=np.float64) for t in tras])
    return {prefix + '_len': lens, prefix + '_is32': is32, prefix + '_wl': wl, prefix + '_tr': tr}


def unpack_curves(d, prefix):
    lens, is32 = d[prefix + '_len'], d[prefix + '_is32']
    off = np.concatenate([[0], np.cumsum(lens)])
    wls, tras = [], []
    for k in range(len(lens)):
        w = d[prefix + '_wl'][off[k]:off[k + 1]]
        t = d[prefix + '_tr'][off[k]:off[k + 1]]
        if is32[k]:
            w, t = w.astype(np.float32), t.astype(np.float32)
        wls.append(list(w)); tras.append(list(t))
    return wls, tras


def problem_from_fixture(path):
    """Rebuild the Problem of a golden fixture: committed observed data and passbands + the
    seeded synthetic grid / isochrone / dust tables (regenerated, not stored)."""
    d = np.load(path, allow_pickle=False)
    cw, ct = unpack_curves(d, 'c')
    pw, pt = unpack_curves(d, 'p')
    bw, bt = unpack_curves(d, 'b')
    ctm = [cw, ct, list(d['c_nresel']), list(d['c_cwl'])]
    ptm = [pw, pt, list(d['p_nresel']), list(d['p_cwl'])]
    bands = []
    for k, name in enumerate(d['b_names']):
        b = PhotBand(str(name), np.asarray(bw[k], dtype=np.float64), np.asarray(bt[k], dtype=np.float64), 'photon',
                     str(d['b_system'][k]), float(d['b_vega'][k]), float(d['b_ab'][k]))
        bands.append(b)
    prob = synthetic.make_problem(
        temps=range(int(d['tmin']), int(d['tmax']) + 1, 100), specmin=float(d['specmin']),
        specmax=float(d['specmax']), dlam=0.2, window_um=(float(d['spmin']), float(d['spmax'])),
        res=int(d['res']), data_wl_um=d['data_wl'], data_flux=d['data_flux'], data_err=d['data_err'],
        contrast_names=[str(s) for s in d['cfilt']], phot_bands=bands, ctm=ctm, ptm=ptm,
        pfilt_names=[str(s) for s in d['pfilt']], cmag=list(d['cmag']), cerr=list(d['cerr']),
        pmag=list(d['pmag']), perr=list(d['perr']), plx=float(d['plx']), plx_err=float(d['plx_err']),
        seed=int(d['seed']), rad_prior=bool(d['rad_prior']), dist_fit=bool(d['dist_fit']))
    prob.tmi, prob.tma = float(d['tmi']), float(d['tma'])
    return prob
