"""``python -m mcmc_spec_b200 -f paramfile -o True -e True`` -- the reference's command line.

Mirror of ``main`` (/root/reference/mft6.py:3450-3708; ``README.md:1``: ``python mft6.py -f [param file]
-o [True/False] -e [True/False]``) for the stages this package implements: parameter file ->
MIST table (:3487-3490) -> observed spectrum (:3492-3509) -> model grid (``spec_interpolator``, :3512,
resampled and broadened on the GPU) -> filter tables, photometry, priors (:3523-3597) -> optimiser
(``-o``, :3655-3664) -> sampler (``-e``, :3666-3690).  The plotting calls (``plot_fit``, ``plot_results``) are
out of scope; the posterior-predictive files they would leave are written when ``--predictive`` is given.

What the reference reads from fixed relative paths or from the network is named explicitly here:

  --models-dir  the BT-Settl directory (reference: ``BT-Settl_M-0.0a+0.0/``, mft6.py:324)
  --mist        the MIST isochrone table (reference: ``mist_2mass_old.cmd``, mft6.py:3487)
  --dust        an .npz with ``dm_edges`` [n] and ``samples`` [n_samples, n]: the Bayestar reddening samples along
                the target's sightline, i.e. what ``bayestar(SkyCoord(ra, dec), mode='samples')`` returns once
                (the reference queries the map inside every posterior call, mft6.py:1233-1234)
  --bps, --vega the passband directory and Vega spectrum (reference: ``bps/``, ``Data/vegaspec.fits``)
  --spectrum    overrides the parameter file's ``filename``

There is no CPU path: the run needs libmcmcspec.so and a CUDA device.
"""
from __future__ import annotations

import getopt
import os
import sys
import time

import numpy as np


def load_mist(path):
    """mft6.py:3487-3490: the table as text, Teff and luminosity columns exponentiated."""
    matrix = np.genfromtxt(path, autostrip=True)
    matrix[:, 4] = 10.0 ** matrix[:, 4]
    matrix[:, 6] = 10.0 ** matrix[:, 6]
    return matrix


def _truthy(s):
    return str(s).strip().lower() in ('true', 't', '1', 'yes')


def main(argv=None):
    argv = sys.argv if argv is None else argv
    opts, _ = getopt.getopt(argv[1:], 'f:o:e:', ['file=', 'optimize=', 'emcee=', 'models-dir=', 'mist=', 'dust=', 'bps=',
                                                 'vega=', 'spectrum=', 'device=', 'seed=', 'predictive'])
    o = dict(opts)
    parfile = o.get('-f', o.get('--file'))
    if parfile is None:
        raise SystemExit(__doc__)
    do_opt = _truthy(o.get('-o', o.get('--optimize', 'False')))
    do_mc = _truthy(o.get('-e', o.get('--emcee', 'False')))
    for need in ('--models-dir', '--mist', '--dust'):
        if need not in o:
            raise SystemExit(f"{need} is required (see python -m mcmc_spec_b200 with no arguments)")
    device, seed = int(o.get('--device', 0)), int(o.get('--seed', 0))
    root = os.path.dirname(os.path.abspath(parfile))
    bps_dir = o.get('--bps', os.path.join(root, 'bps'))
    vega_path = o.get('--vega', os.path.join(root, 'Data', 'vegaspec.fits'))

    from . import fit, gridbuild, paramfile, passbands
    from .posterior import LogPosterior
    fp = paramfile.parse_param_file(parfile)
    if fp.models != 'btsettl':
        raise SystemExit("only models = btsettl is implemented (the hires/PHOENIX reader is out of scope)")
    if int(fp.nspec) not in (2, 3):
        raise SystemExit("nspec must be 2 (binary) or 3 (triple)")
    spectrum = o.get('--spectrum', fp.filename)
    if not os.path.isabs(spectrum) and not os.path.exists(spectrum):
        spectrum = os.path.join(root, spectrum)
    matrix = load_mist(o['--mist'])
    dust = np.load(o['--dust'])
    vega = passbands.load_vega(vega_path) if os.path.exists(vega_path) else None

    t1 = time.time()
    grid, wl, temps, lgs = gridbuild.spec_interpolator(o['--models-dir'], [fp.spmin * 1e4, fp.spmax * 1e4],
                                                       [fp.tmin, fp.tmax], [4, 5.5], [fp.specmin, fp.specmax],
                                                       resolution=fp.res, device=device)            # mft6.py:3512
    print('time to read in specs:', time.time() - t1)                                             # mft6.py:3514
    prob = fit.assemble_problem(fp, spectrum_path=spectrum, bps_dir=bps_dir, grid=grid, wl=wl, temps=temps, lgs=lgs,
                                matrix=matrix, dust_dm_edges=dust['dm_edges'], dust_samples=dust['samples'], vega=vega,
                                parfile_name=os.path.basename(parfile))
    lnprob = LogPosterior(prob, device=device)
    os.makedirs(fp.dirname, exist_ok=True)
    samples = fit.run_fit(lnprob, fp, fp.dirname, optimize=do_opt, emcee=do_mc, seed=seed, nthin=100)
    if do_opt:
        print('optimization complete')                                                            # mft6.py:3664
    if samples is not None and '--predictive' in o and int(fp.nspec) == 2:      # the sweep is defined for binaries
        from . import predictive
        sweep = predictive.ComponentSweep(prob, predictive.kepler_curve(fp.res, bps_dir),
                                          predictive.gaia_g_band(fp.res, bps_dir, vega), device=device, share_grid_with=lnprob)
        predictive.write_sweep(fp.dirname, predictive.kepler_sweep(sweep, samples, rng=np.random.default_rng(seed)))
        sweep.close()
    lnprob.close()
    return 0


if __name__ == '__main__':
    sys.exit(main())
