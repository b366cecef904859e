"""Problem -> oracle glue (tests / smoke / bench CPU baseline only)."""
from types import SimpleNamespace

from oracle import mft6_oracle as mo
from oracle import thirdparty as tp


def to_oracle(prob):
    """Bind a mcmc_spec_b200.Problem to oracle.mft6_oracle.OracleProblem (2-D grids)."""
    names = [b.name for b in prob.photbands]
    photlib = {b.name: tp.LibraryFilter(b.name, b.wavelength, b.transmit, b.dtype, b.vega_zero_flux)
               for b in prob.photbands}
    ns = SimpleNamespace(
        fr=prob.fr, nspec=prob.nspec, data=prob.data, err=prob.err, specs=prob.specs, ctm=prob.ctm,
        ptm=prob.ptm, tmi=prob.tmi, tma=prob.tma, matrix=prob.matrix, temps=list(prob.temps),
        lgs=list(prob.lgs), photlib=photlib,
        dustmap=tp.BayestarSightline(prob.dust_dm_edges, prob.dust_samples), prior=prob.prior,
        av=prob.av, dist_fit=prob.dist_fit, rad_prior=prob.rad_prior, fs=names)
    return mo.OracleProblem(ns)
