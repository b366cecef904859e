"""Oracle-side backend for mcmc_spec_b200.optimizer.fit_spec_batch (TEST code): fit_spec's chi-squares
through oracle.mft6_oracle, one walker at a time, as the reference evaluates them."""
import numpy as np

from oracle import mft6_oracle as mo


class OracleChisq:
    def __init__(self, prob, orc):
        self.p, self.o = prob, orc.p
        self.wl_aa = np.asarray(prob.data[0]) * 1e4
        self.flux = np.asarray(prob.data[1])
        self.reg = [min(prob.data[0]), max(prob.data[0])]
        self.sets = None

    def _common(self):
        o = self.o
        return dict(ctm=o.ctm, ptm=o.ptm, tmi=o.tmi, tma=o.tma, matrix=o.matrix, temps=o.temps, lgs=o.lgs,
                    photlib=o.photlib, dustmap=o.dustmap, nspec=self.p.nspec, dist_fit=o.dist_fit, rad_prior=o.rad_prior, fs=o.fs,
                    nospec=bool(getattr(self.p, 'nospec', False)))

    def start(self, theta0):
        o = self.o
        plx, plx_err = o.prior[-2], o.prior[-1]
        cs, sets = [], []
        for th in np.asarray(theta0):
            ns = self.p.nspec
            c, f = mo.fit_spec_init(self.wl_aa, self.flux, self.p.err, self.reg, list(th[:ns]), [th[ns], 0, 0],
                                    list(th[ns + 1:2 * ns + 1]), o.fr, o.specs, [th[2 * ns + 1], plx, plx_err],
                                    **self._common())
            cs.append(c); sets.append(f)
        self.sets = np.array(sets)
        return np.array(cs), self.sets

    def test(self, theta, walk_index, rad_sigma):
        o = self.o
        plx, plx_err = o.prior[-2], o.prior[-1]
        out = []
        for th, k, rs in zip(np.asarray(theta), walk_index, rad_sigma):
            ns = self.p.nspec
            var_par = [np.array(th[:ns]), th[ns], np.array(th[ns + 1:2 * ns + 1]), th[2 * ns + 1]]
            out.append(mo.fit_spec_test_cs(var_par, self.wl_aa, self.sets[k], self.p.err, self.reg, o.fr, o.specs,
                                           plx, plx_err, rs, **self._common()))
        return np.array(out, dtype=np.float64)
