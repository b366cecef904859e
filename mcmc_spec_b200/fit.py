"""Fit drivers around the device log-posterior: the callers either side of the hot path.

Host-side mirror of the parts of ``main`` / ``run_emcee`` (/root/reference/mft6.py) that feed
and consume ``logposterior``:

  * ``assemble_problem``  -- main's table assembly: observed spectrum (mft6.py:3492-3509), KIC->SDSS
    photometry (:3553-3562), ``fr`` (:3575), ``ctm``/``ptm``/``tmi``/``tma`` (:3577-3597), the prior
    vector handed to the sampler (:3689), bound to a model grid the caller supplies (the BT-Settl
    directory is not in the reference repository; ``gridbuild.spec_interpolator`` reads one when
    present, ``synthetic.make_native_grid`` stands in otherwise).
  * ``run_emcee``         -- the commented sampling block (:1490-1529): burn-in with walker dumps
    every ``nthin`` steps to ``{fname}_{n}_burnin.txt``, reset, production with dumps to
    ``{fname}_{n}_results.txt``, mean integrated autocorrelation time appended to
    ``{fname}_autocorr.txt``, the ``acl * 50 < n`` and ``|d acl| / acl < 0.1`` convergence test,
    the flat chain written to ``samples.txt`` in ``sampler.chain.reshape(-1, ndim)`` order.
  * ``integrated_autocorr_time`` -- emcee's ``get_autocorr_time(quiet=True)`` restated (emcee is not
    vendored by the reference and not installed here): FFT autocorrelation function averaged over
    walkers, Sokal's automatic window with c = 5.
  * ``run_fit``           -- main's ``-o True -e True`` flow (:3655-3690): optimiser, best third of
    its walkers as the ensemble, sampler, optional posterior-predictive sweep.

The ensemble lives on the GPU (``ensemble.EnsembleSampler``); these drivers only move thinned
walker dumps and the final chain to the host, in the reference's text formats so that its
plotting code keeps working on the outputs.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import paramfile, passbands
from .problem import Problem


# ------------------------------------------------------------------------------------------ assembly
def assemble_problem(fp, *, spectrum_path, bps_dir, grid, wl, temps, lgs, matrix, dust_dm_edges, dust_samples,
                     vega=None, parfile_name='', nospec=False) -> Problem:
    """``fp``: paramfile.FitParams.  ``grid`` [n_logg, n_teff, P] on ``wl`` [A] is the (already
    window-broadened) model grid -- the ``specs`` dict of mft6.py:3512; ``matrix`` the MIST table with the
    Teff and L columns exponentiated (:3487-3490); ``dust_*`` the Bayestar sightline samples standing for
    the query of :1233-1234; ``vega`` = passbands.load_vega(...) for the Vega zero points."""
    data_wl, dsp, de = paramfile.load_observed_spectrum(spectrum_path, fp.spmin, fp.spmax, fp.mask)
    ctm, ptm, tmi, tma = passbands.assemble_filter_tables(fp.cfilt, fp.pfilt, fp.res, bps_dir)
    bands = passbands.build_library_bands(len(fp.pfilt), fp.res, bps_dir, vega)
    phot = paramfile.convert_photometry(fp.pmag, fp.pfilt, parfile_name)
    fr = [list(fp.cmag), list(fp.cerr), np.array(fp.cfilt), np.array(phot, dtype=float), list(fp.perr),
          np.array(fp.pfilt)]                                                   # mft6.py:3575
    ndim = 2 * fp.nspec + 2
    prior = [*np.zeros(2 * ndim - 2), fp.plx, fp.plx_err]                      # mft6.py:3689
    return Problem(grid=np.asarray(grid), wl=np.asarray(wl, dtype=np.float64), temps=[int(t) for t in temps],
                   lgs=[float(g) for g in lgs], data=[data_wl, dsp], err=de, fr=fr, ctm=ctm, ptm=ptm,
                   tmi=float(tmi), tma=float(tma), matrix=np.asarray(matrix),
                   dust_dm_edges=np.asarray(dust_dm_edges), dust_samples=np.asarray(dust_samples), photbands=bands,
                   prior=prior, nspec=fp.nspec, av=True, dist_fit=fp.dist_fit, rad_prior=fp.rad_prior, nospec=nospec,
                   meta=dict(res=fp.res, window_um=(fp.spmin, fp.spmax)))


# ------------------------------------------------------------------------------------------ autocorrelation
def _next_pow_two(n):
    i = 1
    while i < n:
        i <<= 1
    return i


def autocorr_function_1d(x):
    """emcee.autocorr.function_1d: normalised autocorrelation function of a 1-D series via FFT."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    if x.ndim != 1:
        raise ValueError("invalid dimensions for 1D autocorrelation function")
    n = _next_pow_two(len(x))
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[:len(x)].real
    acf /= acf[0]
    return acf


def _auto_window(taus, c):
    m = np.arange(len(taus)) < c * taus
    if np.any(m):
        return int(np.argmin(m))
    return len(taus) - 1


def integrated_autocorr_time(chain, c=5):
    """emcee ``integrated_time(x, c=5, quiet=True)`` for ``chain`` [nsteps, nwalkers, ndim] -> tau [ndim].

    ``chain`` may be a numpy array or a torch tensor; the autocorrelation functions of all (walker, parameter)
    series are computed by one batched real FFT with torch.fft on the tensor's device -- for the sampler's chain
    that is the GPU, where it already lives (on the host this estimate, repeated every ``nthin`` steps on a
    growing chain, was the most expensive part of ``run_emcee``)."""
    x = chain if torch.is_tensor(chain) else torch.as_tensor(np.asarray(chain, dtype=np.float64))
    x = x.to(torch.float64)
    if x.ndim == 2:
        x = x[:, :, None]
    n_t, n_w, n_d = x.shape
    n = _next_pow_two(n_t)
    series = x.reshape(n_t, n_w * n_d).T.contiguous()                      # [n_w * n_d, n_t], time axis contiguous
    series = series - series.mean(dim=1, keepdim=True)
    f = torch.fft.rfft(series, n=2 * n, dim=1)
    acf = torch.fft.irfft(f * torch.conj(f), n=2 * n, dim=1)[:, :n_t]       # function_1d of every series
    acf = acf / acf[:, :1]
    fm = acf.reshape(n_w, n_d, n_t).mean(dim=0)                            # mean over walkers, [n_d, n_t]
    taus_all = (2.0 * torch.cumsum(fm, dim=1) - 1.0).cpu().numpy()
    tau = np.empty(n_d)
    for d in range(n_d):
        taus = taus_all[d]
        tau[d] = taus[_auto_window(taus, c)]
    return tau


# ------------------------------------------------------------------------------------------ sampler driver
def _dump(path, coords):
    with open(path, 'ab') as f:                                               # mft6.py:1494-1497
        f.write(b"\n")
        np.savetxt(f, coords)


def run_emcee(dirname, fname, sampler, p0, nsteps, nburn, nthin=10, check_convergence=True, log=None):
    """The sampling block of ``run_emcee`` (mft6.py:1490-1529) on an ``ensemble.EnsembleSampler``.

    Returns ``samples`` [n_kept_steps * nwalkers, ndim] -- ``sampler.chain.reshape(-1, ndim)`` of emcee, i.e.
    walker-major -- after writing it to ``{dirname}/samples.txt``.  The chain is kept on the sampler's device
    (nsteps x nwalkers x ndim doubles) and copied to the host once per ``nthin`` steps for the autocorrelation
    estimate, as the reference does with ``get_autocorr_time``.
    """
    os.makedirs(dirname, exist_ok=True)
    nw, ndim = sampler.nwalkers, sampler.ndim
    sampler.initialize(p0)
    for n in range(nburn):                                                    # sampler.sample(pos, iterations=nburn)
        sampler.step()
        if n % nthin == 0:
            _dump(os.path.join(dirname, '{}_{}_burnin.txt'.format(fname, n)), sampler.coords.detach().cpu().numpy())
    # get_last_sample(); reset(): keep the positions, forget the statistics
    sampler.naccepted.zero_()
    chain = torch.empty((nsteps, nw, ndim), dtype=torch.float64, device=sampler.device)
    old_acl = np.inf
    kept = 0
    acf_path = os.path.join(dirname, '{}_autocorr.txt'.format(fname))
    for n in range(nsteps):
        sampler.step()
        chain[n].copy_(sampler.coords)
        kept = n + 1
        if n % nthin == 0:
            _dump(os.path.join(dirname, '{}_{}_results.txt'.format(fname, n)), chain[n].cpu().numpy())
            with np.errstate(invalid='ignore', divide='ignore'):
                acl = integrated_autocorr_time(chain[:kept])               # sampler.get_autocorr_time(quiet=True)
            macl = np.mean(acl)
            with open(acf_path, 'a') as f:
                f.write(str(macl) + '\n')
            if check_convergence and not np.isnan(macl):                      # mft6.py:1518-1522
                converged = np.all(acl * 50 < n)
                with np.errstate(invalid='ignore'):
                    converged &= np.all((np.abs(old_acl - acl) / acl) < 0.1)
                if converged:
                    break
            old_acl = acl
    steps_since_reset = max(kept, 1)
    af = sampler.naccepted.to(torch.float64).cpu().numpy() / steps_since_reset
    msg = "Mean acceptance fraction: {0:.3f}".format(float(np.mean(af)))      # mft6.py:1525
    (log or print)(msg)
    samples = chain[:kept].permute(1, 0, 2).reshape(-1, ndim).cpu().numpy()   # sampler.chain[:, :, :].reshape((-1, ndim))
    np.savetxt(os.path.join(dirname, 'samples.txt'), samples)
    return samples


def run_fit(lnprob, fp, dirname, optimize=True, emcee=True, seed=0, nthin=100, world=None, log=None):
    """main's ``-o`` / ``-e`` stages (mft6.py:3655-3690) on a ``posterior.LogPosterior``.

    optimize: ``optimizer.optimize_fit`` with ``nwalk`` / ``nstep`` of the parameter file, writing
    ``optimize_res.txt`` / ``optimize_cs.txt``.  emcee: re-read those files, start the ensemble from the best
    third (:3668-3674), ``nburn`` + ``nsteps`` iterations with dumps every 100 (:3689).  Returns the samples
    (or None when ``emcee`` is False).
    """
    from . import optimizer
    from .ensemble import CudaBackend, EnsembleSampler
    os.makedirs(dirname, exist_ok=True)
    if optimize:
        optimizer.optimize_fit(lnprob, fp.nwalk, fp.nstep, fp.plx, fp.plx_err, seed=seed, dirname=dirname)
    if not emcee:
        return None
    chisqs = np.genfromtxt(os.path.join(dirname, 'optimize_cs.txt'))
    pars = np.genfromtxt(os.path.join(dirname, 'optimize_res.txt'))
    p0 = optimizer.select_walkers(pars, chisqs)
    if lnprob.problem.nspec == 3:                                           # mft6.py:3676-3677
        p0 = p0 + np.random.default_rng(seed).normal(np.zeros(np.shape(p0)), 0.05 * np.abs(p0), np.shape(p0))
    if len(p0) % 2:
        p0 = p0[:-1]                       # the red/blue split needs an even ensemble (emcee requires nwalkers >= 2 ndim)
    sampler = EnsembleSampler(len(p0), lnprob.ndim, CudaBackend(lnprob.engine), seed=seed)
    return run_emcee(dirname, fp.fname, sampler, p0, fp.nsteps, fp.nburn, nthin=nthin, log=log)
