"""``Problem``: everything one fit hands to the log-posterior, in the reference's own shapes.

This is the bundle that ``main`` assembles in /root/reference/mft6.py:3487-3597 and passes, by
reference, into every ``logposterior`` call (argument list at mft6.py:1491):

    specs (T1)   dict '{T}, {g}' -> row[P] plus 'wl'      -> here ``grid[n_logg, n_teff, P]`` + ``wl``
    fr    (T2)   [cmag, cerr, cfilt, phot, perr, pfilt]
    ctm/ptm (T3) [[wl lists], [T lists], [n_resel], [mean wl]] for contrast / photometric filters
    data, err (T4)  [wl_um, flux/median], err/median
    matrix       MIST table, Teff and L columns already exponentiated (mft6.py:3489-3490)
    dust table   stands for the Bayestar sightline queried at mft6.py:1233-1234
    photbands    stand for the pyphot library filters of mft6.py:766-783

An optional third grid axis ([Fe/H], fixed per fit) is an extension for the 3-D stress
configuration; the reference has only (Teff, log g).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .passbands import PhotBand


@dataclass
class Problem:
    grid: np.ndarray                 # [n_logg, n_teff, P] (or [n_feh, n_logg, n_teff, P]) float64
    wl: np.ndarray                   # [P] model wavelengths, Angstrom
    temps: List[int]                 # sorted Teff nodes
    lgs: List[float]                 # sorted log g nodes
    data: list                       # [wl_um[N], flux[N]]
    err: np.ndarray                  # [N]
    fr: list                         # [cmag, cerr, cfilt, phot, perr, pfilt]
    ctm: list
    ptm: list
    tmi: float
    tma: float
    matrix: np.ndarray               # MIST-like table (col 1 log age, 4 Teff, 5 log g, 6 L/Lsun)
    dust_dm_edges: np.ndarray        # [n_dm]
    dust_samples: np.ndarray         # [n_samples, n_dm]
    photbands: List[PhotBand]
    prior: list                      # mft6.py:3689 layout: [0]*(2*ndim-2) + [plx, plx_err]
    nspec: int = 2
    av: bool = True
    dist_fit: bool = True
    rad_prior: bool = True
    nospec: bool = False                 # True: the posterior of mft6_nospec.py (no spectral term)
    fehs: Optional[List[float]] = None   # 3-D extension: [Fe/H] nodes
    feh: float = 0.0                     # 3-D extension: the fit's fixed metallicity
    feh_free: bool = False               # 3-D extension: [Fe/H] is a walker parameter (trailing column of theta)
    truth: Optional[np.ndarray] = None   # theta the synthetic data were drawn from
    meta: dict = field(default_factory=dict)

    @property
    def ndim(self):
        return (6 if self.nspec == 2 else 8) + int(self.feh_free)

    @property
    def specs(self):
        """The reference's ``specs`` dict as zero-copy views (2-D grids only)."""
        if self.grid.ndim != 3:
            raise ValueError("specs dict exists for (Teff, log g) grids only")
        d = {'{}, {}'.format(t, g): self.grid[ig, it]
             for ig, g in enumerate(self.lgs) for it, t in enumerate(self.temps)}
        d['wl'] = self.wl
        return d

    @property
    def n_data(self):
        return len(self.err)
