"""``LogPosterior``: the drop-in for mft6.py's ``logposterior`` as an emcee ``log_prob_fn``.

Reference boundary (/root/reference/mft6.py:1491):

    emcee.EnsembleSampler(nwalkers, ndim, logposterior, args=[fr, nspec, ndust, data, err,
        broadening, r, specs, ctm, ptm, tmi, tma, vs, tmin, tmax, matrix, ra, dec],
        kwargs={'dust', 'norm', 'prior', 'a', 'models', 'dist_fit', 'rad_prior'})

Here the ``args`` are bound once (``LogPosterior(problem)`` uploads them to the GPU) and the
callable takes only the walker vector(s):

    lnprob = LogPosterior(problem)            # or LogPosterior.from_param_file(...)
    lnprob(p0)            -> float            # scalar form, p0[ndim]   (emcee default)
    lnprob(P)             -> float64[nw]      # vectorised form, P[nw, ndim] (emcee vectorize=True)

Values are -inf outside the prior box (mft6.py:1227-1230) or when the chi-square is NaN
(mft6.py:1202), never NaN.  There is no CPU path: construction raises if libmcmcspec.so or a
CUDA device is missing.
"""
from __future__ import annotations

import numpy as np

from . import native
from .problem import Problem
from .tables import build_tables


class LogPosterior:
    def __init__(self, problem: Problem, device=0, share_grid_with=None, fast=False):
        """``share_grid_with``: another LogPosterior on the same device, built on the same model grid
        and consumed region, whose device copy of the grid is reused (independent ensembles fitting
        different observed spectra, BASELINE config C3)."""
        self.problem = problem
        self.tables = build_tables(problem)
        owner = None if share_grid_with is None else share_grid_with.engine
        self.engine = native.Engine(self.tables, device=device, grid_owner=owner)
        self.ndim = self.engine.ndim
        if fast:
            self.engine.set_fast(True)         # bracket-only evaluation (ms_set_fast)

    def __call__(self, p0):
        p = np.asarray(p0, dtype=np.float64)
        if p.ndim == 1:
            if p.shape[0] != self.ndim:
                # mft6.py:1456-1457 prints and returns None for a malformed p0; here it is an error
                raise ValueError(f"p0 must have {self.ndim} entries")
            return float(self.engine.lnprob_host(p[None, :])[0])
        return self.engine.lnprob_host(p)

    def parts(self, theta):
        """Intermediate quantities of loglikelihood for a batch (parity debugging)."""
        return self.engine.lnprob_parts_host(np.atleast_2d(np.asarray(theta, dtype=np.float64)))

    def model_spectrum(self, theta):
        """Median-scaled model on the observed pixels (``init_cspec`` after mft6.py:1173)."""
        return self.engine.model_spectrum_host(np.atleast_2d(np.asarray(theta, dtype=np.float64)))

    def component_mags(self, theta):
        """Band magnitude of every component through every contrast filter, [nw, n_contrast, nspec]
        (``mag`` of make_composite, mft6.py:729-733): pass the Kepler / Gaia curves as contrast filters to
        get the per-sample component magnitudes of the reference's posterior-predictive sweep
        (mft6.py:2486-2509)."""
        return self.engine.component_mags_host(np.atleast_2d(np.asarray(theta, dtype=np.float64)))

    def close(self):
        self.engine.close()
