"""Model-grid assembly: the host-side mirror of ``spec_interpolator`` (btsettl branch) with the
instrumental broadening on the GPU.

Reference: /root/reference/mft6.py:323-383 -- discover the (Teff, log g) nodes from the file names
``lte{T/100:03d}-{g}-0.0a+0.0.BT-Settl.spec.7.txt`` (:324-340, ``find_model`` :244-253), read
``wavelength flux`` text rows within +-100 A of the requested range (:353-357), resample every
node onto the common grid ``arange(specmin, specmax, 0.2)`` with ``interp1d`` (:343, :369-371),
broaden **only the data window** to the instrument resolution (``broaden`` :124-130 =
``pyasl.instrBroadGaussFast(maxsig=5)`` + the 5/10-sample patch) and splice it between the
untouched blue and red parts (:373-383).

Only the text reading stays on the host; resampling every node onto the common grid
(``resample_rows_kernel``), the banded broadening stencil of the data window (``broaden_rows_kernel``)
and the splice run on the GPU in one ``ms_build_grid`` call (no CPU fallback: it raises if the
library or a GPU is missing).  The BT-Settl files themselves are not in the reference repository,
so tests write small synthetic files of the same format.
"""
from __future__ import annotations

import glob
import os

import numpy as np

from . import native


def discover_nodes(model_dir, trange, lgrange):
    """Node lists in the order the reference builds them (file-name order, first occurrence), plus the
    file of every (T, g) (mft6.py:324-340, :244-253)."""
    files = glob.glob(os.path.join(model_dir, 'lte*'))
    t, l = [], []
    for f in files:
        base = os.path.basename(f)
        nu = int(float(base.split('-')[0].split('e')[1]) * 1e2)
        mu = float(base.split('-')[1])
        if nu not in t and min(trange) <= nu <= max(trange):
            t.append(nu)
        if mu not in l and min(lgrange) <= mu <= max(lgrange):
            l.append(mu)
    return t, l


def node_file(model_dir, temp, logg):
    """``find_model`` for BT-Settl (mft6.py:244-253)."""
    name = 'lte{}-{}-0.0a+0.0.BT-Settl.spec.7.txt'.format(str(int(temp / 1e2)).zfill(3), str(logg))
    hits = glob.glob(os.path.join(model_dir, name))
    if not hits:
        raise FileNotFoundError(name)
    return hits[0]


def read_node(path, specrange):
    """``wavelength flux`` rows within +-100 A of the requested range (mft6.py:353-357).  The reference
    splits every line on single blanks in a Python loop; here the two leading columns are parsed in bulk
    (same values: both go through the C ``strtod``)."""
    lo, hi = min(specrange) - 100, max(specrange) + 100
    tab = np.loadtxt(path, usecols=(0, 1), dtype=np.float64, ndmin=2)
    keep = (tab[:, 0] >= lo) & (tab[:, 0] <= hi)
    return tab[keep, 0], tab[keep, 1]


def gaussian_taps(window_wl, resolution, maxsig=5):
    """Sum-normalised taps of pyasl.broadGaussFast for an equidistant window (mft6.py:128)."""
    dxs = window_wl[1:] - window_wl[:-1]
    if abs(dxs.max() - dxs.min()) > np.mean(dxs) * 1e-6:
        raise ValueError("The wavelength axis must be equidistant")
    fwhm = np.mean(window_wl) / float(resolution)
    sigma = fwhm / (2.0 * np.sqrt(2.0 * np.log(2.0)))
    lx = int(((sigma * maxsig) / dxs[0]) * 2.0) + 1
    if lx % 2 == 0:
        raise ValueError("an even tap count shifts the kernel by half a pixel; not supported by the stencil")
    nx = (np.arange(lx, dtype=int) - sum(divmod(lx, 2)) + 1) * dxs[0]
    e = np.exp(-nx ** 2 / (2.0 * sigma ** 2))
    return e / e.sum()


def broaden_and_splice(rows, wl, window, resolution, device=0):
    """In place: the data window ``min(window) <= wl <= max(window)`` of every row of ``rows``
    [n_rows, P] is broadened on the GPU; everything outside stays at native resolution
    (mft6.py:373-376).  Returns ``rows``."""
    inwin = np.where((wl >= min(window)) & (wl <= max(window)))[0]
    if inwin.size == 0:
        return rows
    taps = gaussian_taps(wl[inwin], resolution)
    native.broaden_rows(rows, int(inwin[0]), int(inwin.size), taps, device=device)
    return rows


def spec_interpolator(model_dir, w, trange, lgrange, specrange, resolution=10000, device=0):
    """Mirror of ``spec_interpolator(..., models='btsettl')`` (mft6.py:323-383).

    ``w``: data window [A]; returns (grid [n_logg, n_teff, P], wl [P], temps sorted, lgs sorted) -- the
    arrays behind the reference's ``specs`` dict ('{T}, {g}' -> row, 'wl' -> wl).
    """
    t, l = discover_nodes(model_dir, trange, lgrange)
    temps, lgs = sorted(t), sorted(l)
    wl = np.arange(min(specrange), max(specrange), 0.2)               # mft6.py:343
    node_wl, node_fl = [], []
    for g in lgs:
        for temp in temps:
            spold, spec1 = read_node(node_file(model_dir, temp, g), specrange)
            o = np.argsort(spold, kind='mergesort')                    # interp1d(assume_sorted=False)
            node_wl.append(spold[o]); node_fl.append(spec1[o])
    inwin = np.where((wl >= min(w)) & (wl <= max(w)))[0]
    if inwin.size:
        taps = gaussian_taps(wl[inwin], resolution)
        rows = native.build_grid(node_wl, node_fl, wl, int(inwin[0]), int(inwin.size), taps, device=device)
    else:
        rows = native.build_grid(node_wl, node_fl, wl, device=device)
    return rows.reshape(len(lgs), len(temps), wl.size), wl, temps, lgs
