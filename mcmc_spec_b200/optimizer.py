"""Pre-MCMC optimiser stage: ``optimize_fit`` / ``fit_spec`` of mft6.py on the GPU likelihood.

Reference: /root/reference/mft6.py:1686-1765 (``optimize_fit``: ``nwalk`` random starting points,
one ``fit_spec`` per start on a 15-process pool, results appended to ``optimize_res.txt`` /
``optimize_cs.txt``) and :856-1137 (``fit_spec``: a greedy random walk -- propose all parameters at
once from Gaussians around the best point, keep a proposal only if its chi-square is lower, restart
the step counter on success, shrink the step sizes after ``steps/2`` unsuccessful tries, stop after
``steps`` tries without improvement or ``50*steps`` calls).

Here the ``nwalk`` walks advance in lock step: every round, each active walk makes one proposal and
all in-box proposals are evaluated in ONE batch by ``ms_optimizer_chisq`` (fit_spec's ``test_cs``,
mft6.py:995-1046).  The walk logic -- box test (:984), counters, step-size schedule (:971-976), the
fix-up loops that only burn ``total_n`` (:1069-1103) -- is kept line for line; the reference's
process pool and its unseeded ``RandomState`` are replaced by one seeded generator.

``backend`` supplies two calls (the product one is ``GpuChisq``; tests plug in the oracle):
    start(theta0 [nw, ndim]) -> (init_cs [nw], flux_sets [nw, N])      mft6.py:859-930
    test(theta [m, ndim], walk_index [m], rad_sigma [m, nspec]) -> test_cs [m]

Binaries (ndim = 6: T1, T2, Av, R1, R2/R1, plx) and triples (ndim = 8: T1, T2, T3, Av, R1, R2/R1, R3/R1, plx;
the ``nspec == 3`` branches at mft6.py:920-922, :952-976, :984-990, :1044-1050, :1718-1741) share the code:
``ns = (ndim - 2) // 2``.
"""
from __future__ import annotations

import os

import numpy as np


class GpuChisq:
    """fit_spec's two chi-square evaluations through libmcmcspec.so."""

    def __init__(self, lnprob):
        self.lp = lnprob
        self.engine = lnprob.engine
        self.wl_aa = np.asarray(lnprob.problem.data[0], dtype=np.float64) * 1e4     # mft6.py:860
        self.flux = np.asarray(lnprob.problem.data[1], dtype=np.float64)

    def start(self, theta0):
        theta0 = np.asarray(theta0, dtype=np.float64)
        bare = theta0.copy()
        ns = (theta0.shape[1] - 2) // 2
        bare[:, ns] = 0.0                                  # the initial model is not extincted (mft6.py:880)
        if getattr(self.lp.problem, 'nospec', False):
            # mft6_nospec.py: no spectral term, hence no renormalised spectrum (its fit_spec :885-895 commented out)
            sets = np.tile(self.flux, (theta0.shape[0], 1))
        else:
            model = self.engine.model_spectrum_host(bare)  # interp1d + median scaling, mft6.py:884-888
            sets = np.empty_like(model)
            for k in range(model.shape[0]):                # norm_spec, once per walk (mft6.py:889)
                frac = self.flux / model[k]
                sets[k] = self.flux / np.polynomial.Polynomial.fit(self.wl_aa, frac, deg=2)(self.wl_aa)
        self.engine.set_optimizer_data(sets)
        idx = np.arange(theta0.shape[0], dtype=np.int32)
        rad_sigma = 0.05 * theta0[:, ns + 1:2 * ns + 1]     # mft6.py:916
        # the starting model stays scaled to the median of the original spectrum (mft6.py:888 precedes :889)
        init_cs = self.engine.optimizer_chisq(bare, idx, rad_sigma, av_prior=theta0[:, ns],
                                              ref_median=float(np.median(self.flux)))
        return init_cs, sets

    def test(self, theta, walk_index, rad_sigma):
        return self.engine.optimizer_chisq(theta, walk_index, rad_sigma)


def draw_starts(nwalk, tmin, tmax, plx, plx_err, rng, nspec=2):
    """Random starting points of optimize_fit (mft6.py:1712-1741)."""
    rmin, rmax = 0.05, 1.0
    t1 = rng.uniform(tmin, tmax, nwalk)
    t2 = np.array([rng.uniform(tmin, tt) for tt in t1])
    temps = [t1, t2]
    if nspec == 3:
        temps.append(np.array([rng.uniform(tmin, tt) for tt in t2]))
    e1 = rng.uniform(0.1, 0.5, nwalk)
    rg1 = rng.uniform(rmin, rmax, nwalk)
    rg2 = np.array([rng.uniform(rmin, r) / r for r in rg1])
    rads = [rg1, rg2]
    if nspec == 3:
        rads.append(np.array([rng.uniform(rmin, r) / r for r in rg2]))          # mft6.py:1732-1736 (sic: drawn below rg2)
    dist = np.abs(rng.normal(plx, plx_err, nwalk))
    return np.column_stack(temps + [e1] + rads + [dist])


def _in_box(v, tlim):
    """mft6.py:984-985 (the test reads only the first two radius entries, for triples too)."""
    ns = (len(v) - 2) // 2
    return (all(tlim[0] < t < tlim[1] for t in v[:ns]) and 0 <= v[ns] and 0.05 <= v[ns + 1] <= 1.5
            and 0.05 < v[ns + 2] < 1 and 1 / 10 > v[2 * ns + 1] > 1 / 3000)


def _fixup_calls(v, tlim, total_n, cap):
    """The out-of-box branch (mft6.py:1069-1103): the proposal is nudged back towards the box and thrown
    away; all that survives is the number of calls it is charged.  Returns the new ``total_n``."""
    v = v.copy()
    ns = (len(v) - 2) // 2
    temps, rad = v[:ns], v[ns + 1:2 * ns + 1]                      # views
    ip = 2 * ns + 1
    total_n += 1
    while np.any(temps < tlim[0]) and total_n < cap:
        total_n += 1
        temps[temps < tlim[0]] += 100
    while np.any(temps > tlim[1]) and total_n < cap:
        total_n += 1
        temps[temps > tlim[1]] -= 100
    while v[0] < v[1] and total_n < cap:
        total_n += 1
        v[1] -= 100
    while v[ns] < 0 and total_n < cap:
        total_n += 1
        v[ns] += 0.1
    while np.any(rad < 0.05) and total_n < cap:
        total_n += 1
        rad[rad < 0.05] += 0.01
    while v[ip] > 1 / 100 and total_n < cap:
        total_n += 1
        v[ip] -= 0.01 * abs(v[ip])
    while v[ip] < 1 / 3000 and total_n < cap:
        total_n += 1
        v[ip] += 0.01 * abs(v[ip])
    return total_n


def fit_spec_batch(backend, theta0, tlim, steps, rng, record=False):
    """``fit_spec`` for all starting points at once.  Returns (best theta [nw, 6], best chi-square [nw],
    history) where history[k] = list of (theta kept, chi kept, test_cs) per evaluated proposal of walk k
    (only when ``record``; the content of the reference's params{n}.txt / chisq{n}.txt)."""
    theta0 = np.asarray(theta0, dtype=np.float64)
    nw = theta0.shape[0]
    init_cs, _ = backend.start(theta0)
    gi = theta0.copy()
    chi = np.asarray(init_cs, dtype=np.float64).copy()
    ndim = theta0.shape[1]
    ns = (ndim - 2) // 2
    ir, ip = slice(ns + 1, 2 * ns + 1), 2 * ns + 1
    rad0 = theta0[:, ir].copy()                                 # rad_guess: step sizes hang on the *initial* radii
    dist0 = theta0[:, ip].copy()                                # (mft6.py:955, :972)
    d_coarse, d_fine = (0.02, 0.005) if ns == 2 else (0.05, 0.01)      # mft6.py:953-955, :971-974
    n = np.zeros(nw)
    total_n = np.zeros(nw, dtype=np.int64)
    cap = 50 * steps
    history = [[] for _ in range(nw)]
    while True:
        active = np.where((n < steps) & (total_n < cap))[0]
        if active.size == 0:
            break
        fine = n[active] > steps / 2                               # mft6.py:971
        si = np.empty((active.size, ndim))
        si[:, 0:ns] = np.where(fine, 20.0, 250.0)[:, None]
        si[:, ns] = np.where(fine, 0.01, 0.05)
        si[:, ir] = np.where(fine, 0.05, 0.1)[:, None] * rad0[active]
        si[:, ip] = np.where(fine, d_fine, d_coarse) * dist0[active]
        var = gi[active] + si * rng.standard_normal((active.size, ndim))   # make_varied_param, mft6.py:211-228
        ok = np.array([_in_box(v, tlim) for v in var])
        for k, a in enumerate(active):
            if not ok[k]:
                total_n[a] = _fixup_calls(var[k], tlim, int(total_n[a]), cap)
        if ns == 3:
            # mft6.py:987-988: the tertiary ratio is forced below the secondary's
            for k in np.flatnonzero(ok):
                while var[k, ns + 3] >= var[k, ns + 2] or var[k, ns + 3] < 0:
                    var[k, ns + 3] = var[k, ns + 2] * 0.9
        inbox = active[ok]
        if inbox.size:
            total_n[inbox] += 1
            n[inbox] += 1
            test_cs = backend.test(var[ok], inbox.astype(np.int32), si[ok][:, ir])   # mft6.py:995-1046
            better = test_cs < chi[inbox]                                               # NaN / +inf never win
            won = inbox[better]
            gi[won] = var[ok][better]
            chi[won] = test_cs[better]
            n[won] = np.where(n[won] > steps / 2, steps / 2 + 1, 0)                    # mft6.py:1052-1058
            if record:
                for j, a in enumerate(inbox):
                    history[a].append((gi[a].copy(), float(chi[a]), float(test_cs[j])))
    return gi, chi, history


def optimize_fit(lnprob, nwalk, nstep, plx, plx_err, seed=None, dirname=None, backend=None, record=None):
    """Mirror of ``optimize_fit`` (mft6.py:1686-1765) for binaries and triples with a distance.

    ``lnprob``: a LogPosterior (supplies the problem and the GPU engine).  Writes the reference's
    result files into ``dirname`` when given: optimize_res.txt (one best-fit vector per line),
    optimize_cs.txt, and per walk params{n}.txt / chisq{n}.txt.  Returns (params [nwalk, ndim], chisq [nwalk]).
    """
    prob = lnprob.problem
    rng = np.random.default_rng(seed)
    tmin, tmax = float(min(prob.temps)), float(max(prob.temps))
    theta0 = draw_starts(nwalk, tmin, tmax, plx, plx_err, rng, nspec=prob.nspec)
    row = ' '.join(['{}'] * theta0.shape[1]) + '\n'
    backend = backend or GpuChisq(lnprob)
    record = bool(dirname) if record is None else record
    gi, chi, history = fit_spec_batch(backend, theta0, (tmin, tmax), nstep, rng, record=record)
    if dirname:
        os.makedirs(dirname, exist_ok=True)
        for k in range(nwalk):
            with open(os.path.join(dirname, f'params{k}.txt'), 'a') as f:       # mft6.py:1107-1110
                for th, _, _ in history[k]:
                    f.write(row.format(*th))
            with open(os.path.join(dirname, f'chisq{k}.txt'), 'a') as f:        # mft6.py:1112-1115
                for _, c, t in history[k]:
                    f.write('{} {}\n'.format(c, t))
        with open(os.path.join(dirname, 'optimize_res.txt'), 'a') as f:         # mft6.py:1760-1761
            for th in gi:
                f.write(row.format(*th))
        with open(os.path.join(dirname, 'optimize_cs.txt'), 'a') as f:          # mft6.py:1762-1763
            for c in chi:
                f.write(str(c) + '\n')
    return gi, chi


def select_walkers(params, chisqs):
    """Best third of the optimiser output becomes the MCMC ensemble (mft6.py:3668-3674)."""
    order = np.argsort(chisqs, kind='stable')[:int(len(chisqs) * 1 / 3)]
    return np.asarray(params)[order]
