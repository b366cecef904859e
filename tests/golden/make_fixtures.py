"""Generate tests/golden/*.npz from the reference's own data files.  Run in the BUILD container:

    python tests/golden/make_fixtures.py            # needs /root/reference (read-only)

The reference ships no golden vectors for this path and cannot be imported (see oracle/), so the
fixtures pin what *can* be pinned offline:

  * the observed spectra (root koi2298_HET.txt -> config C2, Data/synth_spec_3850_3425.txt -> C1)
    after the reference's own preparation (telluric mask, window cut, median normalisation;
    mft6.py:3492-3509), the contrast/photometry vectors of param_koi2298.txt after the KIC->SDSS
    conversion (mft6.py:3553-3562), and the passband tables get_transmission builds from bps/
    (mft6.py:565-649, :3577-3597) -- all produced by this repo's mirrors of those functions;
  * the oracle's lnprob and intermediate quantities at 64 fixed theta (edge cases included) on
    the seeded synthetic grid (seed 2298), so that both the oracle (CPU suite) and the CUDA path
    (GPU suite) are held to the same committed numbers.

/root/reference does not exist on the GPU box: tests read only the .npz files written here.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from mcmc_spec_b200 import paramfile, passbands, synthetic  # noqa: E402
from golden_io import problem_from_fixture, pack_curves       # noqa: E402
from oracle_adapter import to_oracle                          # noqa: E402

REF = '/root/reference'


def build_inputs(spectrum_path, fp):
    data_wl, dsp, de = paramfile.load_observed_spectrum(spectrum_path, fp.spmin, fp.spmax, fp.mask)
    bps = os.path.join(REF, 'bps')
    ctm, ptm, tmi, tma = passbands.assemble_filter_tables(fp.cfilt, fp.pfilt, fp.res, bps)
    vega = passbands.load_vega(os.path.join(REF, 'Data', 'vegaspec.fits'))
    bands = passbands.build_library_bands(len(fp.pfilt), fp.res, bps, vega)
    phot = paramfile.convert_photometry(fp.pmag, fp.pfilt, 'param_koi2298.txt')
    d = dict(data_wl=data_wl, data_flux=dsp, data_err=de, tmi=float(tmi), tma=float(tma),
             cmag=np.array(fp.cmag), cerr=np.array(fp.cerr), pmag=phot, perr=np.array(fp.perr),
             cfilt=np.array(fp.cfilt), pfilt=np.array(fp.pfilt), plx=fp.plx, plx_err=fp.plx_err,
             res=fp.res, spmin=fp.spmin, spmax=fp.spmax, tmin=fp.tmin, tmax=fp.tmax,
             specmin=fp.specmin, specmax=fp.specmax, rad_prior=fp.rad_prior, dist_fit=fp.dist_fit)
    d.update(pack_curves('c', ctm[0], ctm[1]))
    d.update(pack_curves('p', ptm[0], ptm[1]))
    d['c_nresel'] = np.array(ctm[2], dtype=np.float64); d['p_nresel'] = np.array(ptm[2], dtype=np.float64)
    # mean wavelengths keep their native precision (float32 for the FITS-backed 2MASS J/H curves)
    d['c_cwl'] = np.array([float(c) for c in ctm[3]]); d['p_cwl'] = np.array([float(c) for c in ptm[3]])
    d.update(pack_curves('b', [b.wavelength for b in bands], [b.transmit for b in bands]))
    d['b_names'] = np.array([b.name for b in bands]); d['b_system'] = np.array([b.system for b in bands])
    d['b_vega'] = np.array([b.vega_zero_flux for b in bands]); d['b_ab'] = np.array([b.ab_zero_flux for b in bands])
    return d


def main():
    fp = paramfile.parse_param_file(os.path.join(REF, 'param_koi2298.txt'))
    cases = {
        'c2_koi2298': os.path.join(REF, 'koi2298_HET.txt'),         # param file says Data/..., file ships at the root
        'c1_synth_3850_3425': os.path.join(REF, 'Data', 'synth_spec_3850_3425.txt'),
    }
    for name, path in cases.items():
        d = build_inputs(path, fp)
        d['seed'] = 2298
        tmp = os.path.join(HERE, name + '.npz')
        np.savez_compressed(tmp, **d)
        prob = problem_from_fixture(tmp)
        theta = synthetic.sample_prior_box(prob, 64, seed=11)
        # a cluster around a plausible solution, where the likelihood is not prior dominated
        rng = np.random.default_rng(5)
        centre = np.array([3850.0, 3425.0, 0.1, 0.45, 0.6, fp.plx])
        theta[32:] = centre * (1 + 0.02 * rng.standard_normal((32, 6)))
        theta[32:, 2] = np.abs(theta[32:, 2])
        orc = to_oracle(prob)
        lnp = np.empty(64); chi2 = np.full(64, np.nan); lpr = np.full(64, np.nan)
        con = np.full((64, len(fp.cfilt)), np.nan); ph = np.full((64, len(fp.pfilt)), np.nan)
        for k in range(64):
            parts = {}
            lnp[k] = orc.lnprob(theta[k], parts)
            lpr[k] = parts['lp']
            if 'chi2_spec' in parts:
                chi2[k] = parts['chi2_spec']; con[k] = parts['contrast']; ph[k] = parts['phot']
        d.update(theta=theta, oracle_lnprob=lnp, oracle_chi2_spec=chi2, oracle_lnprior=lpr,
                 oracle_contrast=con, oracle_phot=ph)
        np.savez_compressed(tmp, **d)
        print(name, 'N =', d['data_wl'].size, 'finite lnprob:', int(np.isfinite(lnp).sum()), 'of 64;',
              'size', os.path.getsize(tmp) // 1024, 'KiB')


def make_c3():
    """Config C3: the 12 Data/synth_spec_* spectra as independent fits of one grid.  Stores only what
    differs between the fits (the prepared spectra) + oracle values; passbands etc. come from the C1 fixture."""
    import dataclasses
    import glob
    fp = paramfile.parse_param_file(os.path.join(REF, 'param_koi2298.txt'))
    files = sorted(glob.glob(os.path.join(REF, 'Data', 'synth_spec_*.txt')))
    assert len(files) == 12
    base = problem_from_fixture(os.path.join(HERE, 'c1_synth_3850_3425.npz'))
    rng = np.random.default_rng(17)
    names, lens, wl, fl, er, thetas, lnps = [], [], [], [], [], [], []
    for f in files:
        w, d, e = paramfile.load_observed_spectrum(f, fp.spmin, fp.spmax, fp.mask)
        prob = dataclasses.replace(base, data=[w, d], err=e)
        t1, t2 = [float(x) for x in os.path.basename(f)[:-4].split('_')[2:4]]
        centre = np.array([t1, min(t2, 5400.0), 0.1, 0.5, 0.6, fp.plx])
        theta = centre * (1 + 0.015 * rng.standard_normal((6, 6)))
        theta[:, 2] = np.abs(theta[:, 2])
        lnp = to_oracle(prob).lnprob_batch(theta)
        names.append(os.path.basename(f)); lens.append(w.size); wl.append(w); fl.append(d); er.append(e)
        thetas.append(theta); lnps.append(lnp)
        print(names[-1], 'N =', w.size, 'finite', int(np.isfinite(lnp).sum()))
    out = os.path.join(HERE, 'c3_synth_all.npz')
    np.savez_compressed(out, names=np.array(names), lens=np.array(lens), wl=np.concatenate(wl), flux=np.concatenate(fl),
                        err=np.concatenate(er), theta=np.array(thetas), oracle_lnprob=np.array(lnps))
    print('c3 size', os.path.getsize(out) // 1024, 'KiB')


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'c3':
        make_c3()
    else:
        main()
        make_c3()
