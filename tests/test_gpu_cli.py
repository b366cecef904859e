"""GPU: ``python -m mcmc_spec_b200 -f param -o True -e True`` end to end on a synthetic BT-Settl-style directory,
synthetic bps/ files under the reference's file names, a synthetic MIST table and dust sightline."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(tmp):
    from mcmc_spec_b200 import synthetic
    from test_gridbuild import _write_nodes
    from test_paramfile_passbands import _write_bintable
    models = tmp / 'BT-Settl'; os.makedirs(models)
    _write_nodes(str(models), list(range(3400, 3901, 100)), [4.0, 4.5, 5.0, 5.5], wl_lo=5150.0, wl_hi=24200.0, seed=2)
    bps = tmp / 'bps'; os.makedirs(bps); os.makedirs(tmp / 'Data')
    band = lambda c, w, k, n=160: synthetic.synthetic_passband(c, w, n, ripple_seed=70 + k)     # noqa: E731
    for name, (c, w) in {'DSSI_692nm.dat': (6920, 400), 'DSSI_880nm.dat': (8800, 500),
                         'SLOAN_SDSS.rprime_filter.dat': (6200, 1300), 'SLOAN_SDSS.iprime_filter.dat': (7600, 1400),
                         'SLOAN_SDSS.zprime_filter.dat': (9000, 1400), '2MASS_2MASS.Ks.dat': (21500, 3000),
                         'Kepler_Kepler.K.dat': (6600, 4600)}.items():
        np.savetxt(bps / name, np.column_stack(band(c, w, len(name))))
    g = band(6400, 7000, 3)
    np.savetxt(bps / 'gaia_g_pb.txt', np.column_stack([g[0] / 10.0, g[1]]))                        # nm in the file
    for fil, (c, w) in {'j': (12500, 2600), 'h': (16500, 2800)}.items():
        wl, t = band(c, w, ord(fil), 107)
        _write_bintable(bps / f'2mass_{fil}_001_syn.fits', {'WAVELENGTH': wl.astype(np.float32),
                                                           'THROUGHPUT': t.astype(np.float32), 'ERROR': t * 0})
    vw = np.linspace(3000.0, 26000.0, 900)
    _write_bintable(tmp / 'Data' / 'vegaspec.fits', {'WAVELENGTH': vw, 'FLUX': 1e3 * (vw / 5500.0) ** -2.0})
    m = synthetic.make_isochrone_matrix()
    m[:, 4], m[:, 6] = np.log10(m[:, 4]), np.log10(m[:, 6])                                       # the file holds logs
    np.savetxt(tmp / 'mist.cmd', m)
    dm, ds = synthetic.make_dust_table()
    np.savez(tmp / 'dust.npz', dm_edges=dm, samples=ds)
    w = np.linspace(0.60, 0.76, 420)
    f = 1.0 + 0.1 * np.sin(w * 90.0) + 0.01 * np.random.default_rng(0).standard_normal(w.size)
    np.savetxt(tmp / 'spec.txt', np.column_stack([w, 3e-15 * f, 3e-17 * np.ones_like(w)]))
    run = tmp / 'run'
    par = tmp / 'param_test.txt'
    par.write_text('\n'.join([
        '# synthetic system', 'models btsettl', f'dirname {run}', 'fname sys', 'res 1700', 'tmin 3400', 'tmax 3900',
        'specmin 5300', 'specmax 24000', 'mask True', 'rad_prior True', 'filename spec.txt', 'spmin 0.62', 'spmax 0.74',
        'cmag [1.5,1.1]', 'cerr [0.1,0.1]', "cfilt ['692','880']", 'pmag [14.0,13.6,13.4,12.4,11.9,11.7]',
        'perr [0.1,0.1,0.1,0.03,0.03,0.03]', "pfilt ['sdss,r','sdss,i','sdss,z','j','h','k']", 'plx 2.5e-3',
        'plx_err 5e-5', 'dist_fit True', 'av 0.1', 'av_err 0.01', 'ra 288.4', 'dec 45.8', 'nwalk 36', 'nstep 3',
        'nspec 2', 'ndust 0', 'nburn 11', 'nsteps 21', 'real_values [0,0,0,0,0,0,0,0]', '']))
    return par, models, run


def test_command_line_flow(tmp_path, native_lib):
    from mcmc_spec_b200.__main__ import main
    par, models, run = _setup(tmp_path)
    rc = main(['mcmc_spec_b200', '-f', str(par), '-o', 'True', '-e', 'True', '--models-dir', str(models),
               '--mist', str(tmp_path / 'mist.cmd'), '--dust', str(tmp_path / 'dust.npz'), '--seed', '4', '--predictive'])
    assert rc == 0
    res = np.genfromtxt(run / 'optimize_res.txt')
    assert res.shape == (36, 6) and np.genfromtxt(run / 'optimize_cs.txt').shape == (36,)
    samples = np.loadtxt(run / 'samples.txt')
    assert samples.shape == (21 * 12, 6) and np.all(np.isfinite(samples))            # best third of 36 walks
    assert (run / 'sys_0_burnin.txt').exists() and (run / 'sys_0_results.txt').exists() and (run / 'sys_autocorr.txt').exists()
    for name in ('kep_contrast.txt', 'gaia_sec.txt', 'gaia_pri.txt', 'pri_corr.txt', 'sec_corr.txt'):
        assert np.all(np.isfinite(np.loadtxt(run / name)))
    assert np.all((samples[:, 0] >= 3400) & (samples[:, 0] <= 3900))                  # the chain respects the prior box
    # -e alone re-reads the optimiser files (mft6.py:3668)
    assert main(['mcmc_spec_b200', '-f', str(par), '-o', 'False', '-e', 'True', '--models-dir', str(models),
                 '--mist', str(tmp_path / 'mist.cmd'), '--dust', str(tmp_path / 'dust.npz')]) == 0
