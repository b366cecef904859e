"""Fit drivers (run_emcee mirror, autocorrelation time, problem assembly) on CPUs."""
import os

import numpy as np
import pytest

from mcmc_spec_b200 import fit
from mcmc_spec_b200.ensemble import EnsembleSampler
import stretch_reference as sr

REF = '/root/reference'


def test_integrated_autocorr_time_of_ar1():
    """AR(1) with coefficient rho: tau_int = (1 + rho) / (1 - rho)."""
    rng = np.random.default_rng(4)
    rho, n, nw = 0.8, 20000, 16
    x = np.zeros((n, nw, 1))
    e = rng.standard_normal((n, nw))
    for t in range(1, n):
        x[t, :, 0] = rho * x[t - 1, :, 0] + e[t]
    tau = fit.integrated_autocorr_time(x)
    assert abs(tau[0] - 9.0) < 0.9
    acf = fit.autocorr_function_1d(x[:, 0, 0])
    assert acf[0] == 1.0 and abs(acf[1] - rho) < 0.03
    # white noise: tau ~ 1
    assert abs(fit.integrated_autocorr_time(rng.standard_normal((5000, 8, 2)))[1] - 1.0) < 0.15


def test_run_emcee_files_and_chain_layout(tmp_path):
    mu, sig = np.array([1.0, -2.0, 0.5]), np.array([0.5, 2.0, 1.0])
    lnp = lambda th: -0.5 * np.sum(((th - mu) / sig) ** 2, axis=1)   # noqa: E731
    nw, ndim = 20, 3
    s = EnsembleSampler(nw, ndim, sr.NumpyBackend(lnp, ndim), seed=9)
    p0 = mu + 0.1 * np.random.default_rng(1).standard_normal((nw, ndim))
    msgs = []
    samples = fit.run_emcee(str(tmp_path), 'koi', s, p0, nsteps=120, nburn=25, nthin=10, check_convergence=False,
                            log=msgs.append)
    assert samples.shape == (120 * nw, ndim)
    # walker-major (emcee's sampler.chain.reshape(-1, ndim)): the last row is the last walker's final position
    assert np.array_equal(samples[-1], s.coords.numpy()[-1]) and np.array_equal(samples[119], s.coords.numpy()[0])
    assert np.allclose(np.loadtxt(tmp_path / 'samples.txt'), samples)
    for n in (0, 10, 20):
        assert np.loadtxt(tmp_path / f'koi_{n}_burnin.txt').shape == (nw, ndim)
    assert not (tmp_path / 'koi_30_burnin.txt').exists()
    assert np.array_equal(np.loadtxt(tmp_path / 'koi_110_results.txt'), samples.reshape(nw, 120, ndim)[:, 110])
    assert len(open(tmp_path / 'koi_autocorr.txt').read().split()) == 12
    assert msgs and msgs[0].startswith('Mean acceptance fraction: 0.')


def test_run_emcee_stops_when_converged(tmp_path):
    lnp = lambda th: -0.5 * np.sum(th ** 2, axis=1)   # noqa: E731
    s = EnsembleSampler(32, 2, sr.NumpyBackend(lnp, 2), seed=2)
    p0 = np.random.default_rng(0).standard_normal((32, 2))
    samples = fit.run_emcee(str(tmp_path), 'x', s, p0, nsteps=4000, nburn=50, nthin=100, log=lambda m: None)
    kept = samples.shape[0] // 32
    assert kept < 4000 and (kept - 1) % 100 == 0          # broke out at a thinning step (mft6.py:1518-1522)
    tau = fit.integrated_autocorr_time(samples.reshape(32, kept, 2).transpose(1, 0, 2))
    assert np.all(tau * 50 < kept)


@pytest.mark.skipif(not os.path.exists(REF), reason='reference checkout not present (GPU box)')
def test_assemble_problem_from_the_reference_param_file():
    from mcmc_spec_b200 import paramfile, passbands, synthetic
    from mcmc_spec_b200.tables import build_tables
    fp = paramfile.parse_param_file(os.path.join(REF, 'param_koi2298.txt'))
    temps, lgs = list(range(fp.tmin, fp.tmax + 1, 100)), [4.0, 4.5, 5.0, 5.5]
    wl = np.arange(fp.specmin, fp.specmax, 0.2)
    grid = np.ones((len(lgs), len(temps), wl.size))       # table assembly does not look at the values
    dm, ds = synthetic.make_dust_table()
    vega = passbands.load_vega(os.path.join(REF, 'Data', 'vegaspec.fits'))
    prob = fit.assemble_problem(fp, spectrum_path=os.path.join(REF, 'koi2298_HET.txt'), bps_dir=os.path.join(REF, 'bps'),
                                grid=grid, wl=wl, temps=temps, lgs=lgs, matrix=synthetic.make_isochrone_matrix(),
                                dust_dm_edges=dm, dust_samples=ds, vega=vega, parfile_name='param_koi2298.txt')
    assert prob.n_data == 1945 and prob.ndim == 6
    assert prob.prior[-2:] == [fp.plx, fp.plx_err] and len(prob.prior) == 12
    assert list(prob.fr[2]) == fp.cfilt and len(prob.fr[3]) == 6
    t = build_tables(prob)
    assert t.n_reg == 93245 and t.n_data == 1945          # P_reg and N of config C2


def test_command_line_argument_checks(tmp_path):
    """python -m mcmc_spec_b200: usage text without -f, explicit paths for what the reference hard-codes, and the
    MIST loader's exponentiation of the Teff / luminosity columns (mft6.py:3487-3490)."""
    from mcmc_spec_b200.__main__ import load_mist, main
    from mcmc_spec_b200 import synthetic
    with pytest.raises(SystemExit) as e:
        main(['mcmc_spec_b200'])
    assert 'python -m mcmc_spec_b200 -f paramfile' in str(e.value)
    par = tmp_path / 'p.txt'
    par.write_text('models btsettl\n')
    with pytest.raises(SystemExit) as e:
        main(['mcmc_spec_b200', '-f', str(par), '-o', 'True'])
    assert '--models-dir is required' in str(e.value)
    m = synthetic.make_isochrone_matrix()
    logs = m.copy(); logs[:, 4], logs[:, 6] = np.log10(m[:, 4]), np.log10(m[:, 6])
    np.savetxt(tmp_path / 'mist.cmd', logs)
    back = load_mist(str(tmp_path / 'mist.cmd'))
    assert np.allclose(back[:, 4], m[:, 4], rtol=1e-12) and np.allclose(back[:, 6], m[:, 6], rtol=1e-12)
