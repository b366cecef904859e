"""Problem -> oracle glue (tests / smoke / bench CPU baseline only)."""
from types import SimpleNamespace

from oracle import mft6_oracle as mo
from oracle import thirdparty as tp


def _specs_at_feh(prob, feh=None):
    """3-D grids (an extension; the reference has only Teff x log g): the reference's 2-D path runs on
    the grid linearly interpolated to the fit's fixed [Fe/H], with get_spec's own bracket rule
    (nearest node + partner on the far side, mft6.py:439-453) and interp_2_spec (mft6.py:198-209)."""
    feh = prob.feh if feh is None else feh
    z1, z2 = mo._bracket(list(prob.fehs), feh)
    i1, i2 = list(prob.fehs).index(z1), list(prob.fehs).index(z2)
    specs = {'wl': prob.wl}
    for ig, g in enumerate(prob.lgs):
        for it, t in enumerate(prob.temps):
            key = '{}, {}'.format(t, g)
            if i1 == i2:
                specs[key] = prob.grid[i1, ig, it]
            else:
                specs[key] = mo.interp_2_spec(prob.grid[i1, ig, it], prob.grid[i2, ig, it], z1, z2, feh)
    return specs


class FreeFehOracle:
    """Free-metallicity extension: every walker carries its own [Fe/H] (trailing column); the oracle value is
    the reference's 6-parameter posterior on the grid interpolated to that walker's [Fe/H]; outside the nodes
    the walker is outside the prior box (-inf)."""

    def __init__(self, prob):
        self.prob = prob

    def lnprob_batch(self, theta):
        import dataclasses
        import numpy as np
        out = np.empty(len(theta))
        for k, th in enumerate(np.asarray(theta, dtype=float)):
            z = th[-1]
            if not (min(self.prob.fehs) <= z <= max(self.prob.fehs)) or not np.all(np.isfinite(th)):
                out[k] = -np.inf
                continue
            fixed = dataclasses.replace(self.prob, feh=float(z), feh_free=False)
            out[k] = to_oracle(fixed).lnprob(th[:-1])
        return out


def to_oracle(prob):
    """Bind a mcmc_spec_b200.Problem to oracle.mft6_oracle.OracleProblem."""
    if getattr(prob, 'feh_free', False):
        return FreeFehOracle(prob)
    names = [b.name for b in prob.photbands]
    photlib = {b.name: tp.LibraryFilter(b.name, b.wavelength, b.transmit, b.dtype, b.vega_zero_flux)
               for b in prob.photbands}
    ns = SimpleNamespace(
        fr=prob.fr, nspec=prob.nspec, data=prob.data, err=prob.err,
        specs=prob.specs if prob.fehs is None else _specs_at_feh(prob), ctm=prob.ctm,
        ptm=prob.ptm, tmi=prob.tmi, tma=prob.tma, matrix=prob.matrix, temps=list(prob.temps),
        lgs=list(prob.lgs), photlib=photlib,
        dustmap=tp.BayestarSightline(prob.dust_dm_edges, prob.dust_samples), prior=prob.prior,
        av=prob.av, dist_fit=prob.dist_fit, rad_prior=prob.rad_prior, fs=names,
        nospec=getattr(prob, 'nospec', False))
    return mo.OracleProblem(ns)
