"""GPU parity: the CUDA path, called through the C ABI, against the oracle and the golden vectors.

Bar (BASELINE.json north_star): |lnp_gpu - lnp_oracle| <= 1e-9 * |lnp_oracle| per walker, identical
-inf sets.  Nothing here reads /root/reference.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9          # the tolerance north_star states
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _compare(got, want, rtol=RTOL):
    assert not np.any(np.isnan(got))                       # emcee raises on NaN log-prob
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    assert np.all(got[~np.isfinite(want)] == -np.inf)
    f = np.isfinite(want)
    rel = np.abs(got[f] - want[f]) / np.abs(want[f])
    assert rel.max() <= rtol, f"max rel err {rel.max():.3e}"
    return rel.max()


@pytest.fixture(scope='module')
def small_gpu(small_problem, native_lib):
    from mcmc_spec_b200.posterior import LogPosterior
    lp = LogPosterior(small_problem)
    yield lp
    lp.close()


def test_small_problem_2048_walkers(small_problem, small_oracle, small_gpu):
    from mcmc_spec_b200 import synthetic
    theta = synthetic.sample_prior_box(small_problem, 2048, seed=101)
    worst = _compare(small_gpu(theta), small_oracle.lnprob_batch(theta))
    assert worst < 1e-12                                    # in practice rounding level


def test_scalar_and_vector_forms_agree(small_problem, small_gpu):
    from mcmc_spec_b200 import synthetic
    theta = synthetic.sample_prior_box(small_problem, 32, seed=7)
    vec = small_gpu(theta)
    for k in (0, 5, 8, 20):
        s = small_gpu(theta[k])
        assert isinstance(s, float) and (s == vec[k] or (np.isinf(s) and np.isinf(vec[k])))
    with pytest.raises(ValueError):
        small_gpu(np.zeros(5))


def test_parts_match_oracle(small_problem, small_oracle, small_gpu):
    from mcmc_spec_b200 import synthetic
    theta = synthetic.sample_prior_box(small_problem, 64, seed=9, edge_cases=False)
    parts = small_gpu.parts(theta)
    model = small_gpu.model_spectrum(theta[:8])
    for k in range(64):
        ref = {}
        small_oracle.lnprob(theta[k], ref)
        assert parts['lnprior'][k] == pytest.approx(ref['lp'], rel=1e-12)
        assert parts['chi2_spec'][k] == pytest.approx(ref['chi2_spec'], rel=1e-11)
        assert np.allclose(parts['contrast'][k], ref['contrast'], rtol=1e-12)
        assert np.allclose(parts['phot'][k], ref['phot'], rtol=1e-13)
        assert np.allclose(parts['logg'][k], ref['logg'], rtol=1e-14)
        assert parts['status'][k] == 0
        if k < 8:
            assert np.allclose(model[k], ref['model'], rtol=1e-13)


def test_component_magnitudes(small_problem, small_oracle, small_gpu):
    """Per-component band magnitudes (make_composite's ``mag``): their difference is the contrast."""
    from mcmc_spec_b200 import synthetic
    theta = synthetic.sample_prior_box(small_problem, 40, seed=19, edge_cases=False)
    mags = small_gpu.component_mags(theta)
    parts = small_gpu.parts(theta)
    assert mags.shape == (40, 2, 2)
    assert np.allclose(mags[:, :, 1] - mags[:, :, 0], parts['contrast'], rtol=1e-13, atol=1e-13)
    from oracle import mft6_oracle as mo
    o = small_oracle.p
    k = 3
    lg = [mo.get_logg(t, o.matrix) for t in theta[k, :2]]
    w, spec, con, cwl, ph = mo.make_composite(theta[k, :2], lg, theta[k, 3:5], theta[k, 5], o.fr[2], o.fr[5], small_oracle.r,
                                              o.specs, o.ctm, o.ptm, o.tmi, o.tma, o.temps, o.lgs, o.photlib, fs=o.fs)
    assert np.allclose(mags[k, :, 1] - mags[k, :, 0], con, rtol=1e-12)
    bad = theta[:2].copy(); bad[0, 2] = -1.0
    assert np.all(np.isnan(small_gpu.component_mags(bad)[0])) and np.all(np.isfinite(small_gpu.component_mags(bad)[1]))


def test_status_bits_and_domain_failures(small_problem, small_gpu):
    from mcmc_spec_b200 import native
    t = small_problem.truth
    bad = np.array([t, t, t, t, t])
    bad[1, 2] = -0.1            # Av < 0
    bad[2, 0] = 9000.0          # above tmax
    bad[3, 3] = np.nan          # non-finite theta
    bad[4, 5] = 0.3             # parallax > 1/4
    p = small_gpu.parts(bad)
    assert np.isfinite(p['lnprob'][0]) and np.all(p['lnprob'][1:] == -np.inf)
    assert p['status'][1] == native.MS_W_PRIOR_BOUNDS and p['status'][2] == native.MS_W_PRIOR_BOUNDS
    assert p['status'][3] == native.MS_W_NONFINITE and p['status'][4] == native.MS_W_PRIOR_BOUNDS


def test_batch_split_and_permutation_are_bit_identical(small_problem, small_gpu):
    from mcmc_spec_b200 import synthetic
    theta = synthetic.sample_prior_box(small_problem, 600, seed=33)
    small_gpu.engine.set_cluster(1)                         # one CTA per walker whatever the batch size
    full = small_gpu(theta)
    parts = np.concatenate([small_gpu(theta[:77]), small_gpu(theta[77:301]), small_gpu(theta[301:])])
    assert np.array_equal(full, parts)
    perm = np.random.default_rng(0).permutation(600)
    assert np.array_equal(small_gpu(theta[perm]), full[perm])
    assert small_gpu(theta[:0]).shape == (0,)               # empty batch
    small_gpu.engine.set_cluster(0)


def test_paired_walkers_are_bit_identical(small_problem, small_oracle, small_gpu):
    """Large batches run two cell-mates per CTA (ms_set_pairing): a walker's value must not depend on that,
    nor on who its partner is (odd batch size, prior-rejected partners, shuffled order)."""
    from mcmc_spec_b200 import synthetic
    theta = synthetic.sample_prior_box(small_problem, 2 * 1184 + 1, seed=77)
    eng = small_gpu.engine
    eng.set_pairing(False)
    single = small_gpu(theta)
    eng.set_pairing(True)
    paired = small_gpu(theta)
    assert np.array_equal(single, paired)
    assert np.isneginf(paired).any() and np.isfinite(paired).any()
    perm = np.random.default_rng(3).permutation(len(theta))
    assert np.array_equal(small_gpu(theta[perm]), paired[perm])
    # all walkers in one grid cell: every CTA holds a true pair
    t = np.tile(small_problem.truth, (2400, 1))
    t[:, 0] += np.linspace(-20.0, 20.0, 2400); t[:, 3] *= np.linspace(0.9, 1.1, 2400)
    eng.set_pairing(False); a = small_gpu(t)
    eng.set_pairing(True); b = small_gpu(t)
    assert np.array_equal(a, b) and np.isfinite(b).all()
    want = small_oracle.lnprob_batch(t[::300])
    assert np.max(np.abs(b[::300] - want) / np.abs(want)) <= 1e-9


def test_cluster_split_of_a_walker(small_problem, small_oracle, small_gpu):
    """Thread-block clusters of 2/4/8 CTAs per walker (small ensembles): same values to rounding, and the
    automatic choice kicks in for a 50-walker ensemble like the reference's."""
    from mcmc_spec_b200 import synthetic
    theta = synthetic.sample_prior_box(small_problem, 50, seed=71)
    want = small_oracle.lnprob_batch(theta)
    small_gpu.engine.set_cluster(1)
    one = small_gpu(theta)
    for c in (2, 4, 8, 0):
        small_gpu.engine.set_cluster(c)
        got = small_gpu(theta)
        _compare(got, want)
        f = np.isfinite(one)
        assert np.array_equal(np.isfinite(got), f) and np.allclose(got[f], one[f], rtol=1e-13, atol=0)
        assert np.array_equal(got, small_gpu(theta))        # deterministic within a mode
        parts = small_gpu.parts(theta)
        model = small_gpu.model_spectrum(theta[16:20])
        assert np.allclose(np.median(model, axis=1), np.median(small_problem.data[1]), rtol=1e-14)
        assert np.array_equal(parts['lnprob'], got)
    small_gpu.engine.set_cluster(0)


@pytest.mark.parametrize('n_data', [7, 255, 256, 300, 301, 1025])
def test_ragged_and_odd_even_sizes(n_data):
    """median over odd/even N, N not a multiple of the CTA width, N < one warp."""
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    p = synthetic.make_problem_small(n_data=n_data, seed=40 + n_data)
    lp = LogPosterior(p)
    theta = synthetic.sample_prior_box(p, 48, seed=n_data)
    _compare(lp(theta), to_oracle(p).lnprob_batch(theta))
    lp.close()


def test_duplicate_model_values_in_median():
    """Flat model spectrum -> many equal keys in the radix select."""
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    p = synthetic.make_problem_small(n_data=200, seed=77)
    p.grid[...] = 1.0e6                                     # every row flat: model on data pixels is constant x extinction
    lp = LogPosterior(p)
    theta = synthetic.sample_prior_box(p, 32, seed=5)
    theta[:8, 2] = 0.0                                      # Av = 0: exactly constant model
    _compare(lp(theta), to_oracle(p).lnprob_batch(theta))
    lp.close()


@pytest.mark.parametrize('name', ['c2_koi2298', 'c1_synth_3850_3425'])
def test_golden_fixtures_full_size(name, native_lib):
    """BASELINE configs C1/C2: committed observed data + passbands, 135 000-pixel rows, 64 fixed theta."""
    import golden_io
    from mcmc_spec_b200.posterior import LogPosterior
    path = os.path.join(GOLD, name + '.npz')
    fx = np.load(path)
    lp = LogPosterior(golden_io.problem_from_fixture(path))
    _compare(lp(fx['theta']), fx['oracle_lnprob'])
    parts = lp.parts(fx['theta'])
    f = np.isfinite(fx['oracle_lnprob'])
    assert np.allclose(parts['chi2_spec'][f], fx['oracle_chi2_spec'][f], rtol=1e-10)
    assert np.allclose(parts['contrast'][f], fx['oracle_contrast'][f], rtol=1e-11)
    assert np.allclose(parts['phot'][f], fx['oracle_phot'][f], rtol=1e-12)
    lp.close()


def test_c3_twelve_independent_fits_share_one_grid(native_lib):
    """BASELINE config C3: the 12 Data/synth_spec_* spectra fitted as independent problems on one GPU with
    one device copy of the grid; each must equal both the committed oracle values and the stand-alone
    (own-grid) engine bit for bit."""
    import dataclasses
    import golden_io
    from mcmc_spec_b200.posterior import LogPosterior
    fx = np.load(os.path.join(GOLD, 'c3_synth_all.npz'))
    base = golden_io.problem_from_fixture(os.path.join(GOLD, 'c1_synth_3850_3425.npz'))
    off = np.concatenate([[0], np.cumsum(fx['lens'])])
    owner, engines = None, []
    for k in range(12):
        sl = slice(off[k], off[k + 1])
        prob = dataclasses.replace(base, data=[fx['wl'][sl], fx['flux'][sl]], err=fx['err'][sl])
        lp = LogPosterior(prob, share_grid_with=owner)
        owner = owner or lp
        engines.append(lp)
    for k, lp in enumerate(engines):
        got = lp(fx['theta'][k])
        _compare(got, fx['oracle_lnprob'][k])
        if k in (3, 11):                                  # stand-alone engine with its own grid upload
            sl = slice(off[k], off[k + 1])
            solo = LogPosterior(dataclasses.replace(base, data=[fx['wl'][sl], fx['flux'][sl]], err=fx['err'][sl]))
            assert np.array_equal(solo(fx['theta'][k]), got)
            solo.close()
    for lp in reversed(engines):                          # the grid owner goes last
        lp.close()


def test_c4_shape_against_oracle_and_properties(native_lib):
    """BASELINE config C4 (8192 observed pixels, 32 768-pixel rows): oracle on a sample + size-independent
    properties on the full batch."""
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    p = synthetic.make_problem_c4()
    lp = LogPosterior(p)
    theta = synthetic.sample_prior_box(p, 8192, seed=55)
    got = lp(theta)
    _compare(got[:96], to_oracle(p).lnprob_batch(theta[:96]))
    # (i) the median-scaled model has the data's median (mft6.py:1173)
    model = lp.model_spectrum(theta[100:132])
    assert np.allclose(np.median(model, axis=1), np.median(p.data[1]), rtol=1e-14)
    # (ii) spectral chi-square and contrasts do not depend on the overall flux scale: R1 -> 1.3 R1
    a = lp.parts(theta[200:264]); th2 = theta[200:264].copy(); th2[:, 3] *= 1.3
    b = lp.parts(th2)
    ok = (a['status'] == 0) & (b['status'] == 0)
    assert ok.sum() > 20
    assert np.allclose(a['chi2_spec'][ok], b['chi2_spec'][ok], rtol=1e-10)
    assert np.allclose(a['contrast'][ok], b['contrast'][ok], rtol=1e-12, atol=1e-12)
    assert np.allclose(b['phot'][ok] - a['phot'][ok], -5 * np.log10(1.3), rtol=1e-10)
    # (iii) determinism
    assert np.array_equal(lp(theta), got)
    lp.close()


def test_3d_grid_trilinear_16_passbands(native_lib):
    """Shape of BASELINE config C5 (Teff x log g x [Fe/H] grid, 4 contrast + 12 photometric bands) at a size
    the oracle finishes quickly; the oracle runs the reference's 2-D path on the grid interpolated to the fit's
    fixed [Fe/H] (tests/oracle_adapter.py)."""
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    for feh in (0.2, 0.0, -0.5, 0.25):                      # off-node, on-node, on the edge node, midpoint tie
        p = synthetic.make_problem_3d_small(feh=feh)
        lp = LogPosterior(p)
        theta = synthetic.sample_prior_box(p, 96, seed=int(10 * (feh + 1)))
        _compare(lp(theta), to_oracle(p).lnprob_batch(theta))
        lp.close()


def test_bracket_only_mode_equals_streaming(small_problem, small_oracle, native_lib):
    """ms_set_fast: two model pixels per observed pixel + tabulated per-row band fluxes instead of streaming whole
    rows (algebraic note).  Model vector bit-identical to the streaming path, chi-square, contrasts
    and magnitudes to rounding, log-posterior within 1e-12 of it and 1e-9 of the oracle -- for binaries, triples,
    3-D grids, the nospec flavour, a free [Fe/H] and a spectrum longer than shared memory."""
    import dataclasses
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    cases = [(small_problem, synthetic.sample_prior_box(small_problem, 512, seed=41)),
             (dataclasses.replace(small_problem, nospec=True), synthetic.sample_prior_box(small_problem, 128, seed=42))]
    pt = synthetic.make_problem_triple_small()
    cases.append((pt, synthetic.sample_prior_box_triple(pt, 128, seed=43)))
    p3 = synthetic.make_problem_3d_small(feh=0.2)
    cases.append((p3, synthetic.sample_prior_box(p3, 128, seed=44)))
    p3f = synthetic.make_problem_3d_small(feh_free=True)
    cases.append((p3f, synthetic.sample_prior_box(p3f, 128, seed=45)))
    pl = synthetic.make_problem_small(n_pix=140000, dlam=0.03, n_data=30011, window_um=(0.56, 0.94))
    cases.append((pl, synthetic.sample_prior_box(pl, 48, seed=46)))
    for prob, theta in cases:
        lp = LogPosterior(prob)
        a = lp.parts(theta)
        lp.engine.set_fast(True)
        b = lp.parts(theta)
        fast = lp(theta)
        lp.engine.set_fast(False)
        assert np.array_equal(a['status'], b['status'])
        ok = a['status'] == 0
        assert ok.sum() > 20
        if not prob.nospec:
            # same model vector (bit for bit); the norm_spec moments are summed in another order
            assert np.allclose(a['chi2_spec'][ok], b['chi2_spec'][ok], rtol=1e-12, atol=0)
            k = np.flatnonzero(ok)[:4]
            lp.engine.set_fast(True); mf = lp.model_spectrum(theta[k])
            lp.engine.set_fast(False); ms = lp.model_spectrum(theta[k])
            assert np.array_equal(mf, ms)
        assert np.allclose(a['contrast'][ok], b['contrast'][ok], rtol=0, atol=1e-11)
        assert np.allclose(a['phot'][ok], b['phot'][ok], rtol=0, atol=1e-11)
        assert np.array_equal(np.isfinite(a['lnprob']), np.isfinite(b['lnprob']))
        f = np.isfinite(a['lnprob'])
        assert np.max(np.abs(a['lnprob'][f] - b['lnprob'][f]) / np.abs(a['lnprob'][f])) <= 1e-12
        assert np.array_equal(fast, b['lnprob'])
        _compare(fast[:32], to_oracle(prob).lnprob_batch(theta[:32]))
        lp.close()
    lpf = LogPosterior(small_problem, fast=True)                                   # constructor switch
    ref = LogPosterior(small_problem)
    th = cases[0][1][:64]
    f = np.isfinite(ref(th))
    assert np.allclose(lpf(th)[f], ref(th)[f], rtol=1e-12) and np.array_equal(np.isfinite(lpf(th)), f)
    lpf.close(); ref.close()


def test_observed_spectrum_longer_than_shared_memory(native_lib):
    """More observed pixels than the shared-memory resident model vector can hold (~27 000): the per-walker model
    vector moves to a global scratch row; same values, same edge behaviour."""
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    p = synthetic.make_problem_small(n_pix=140000, dlam=0.03, n_data=40001, window_um=(0.56, 0.94))
    lp = LogPosterior(p)
    assert lp.tables.n_data == 40001
    theta = synthetic.sample_prior_box(p, 300, seed=9)
    got = lp(theta)
    _compare(got[:40], to_oracle(p).lnprob_batch(theta[:40]))
    assert np.array_equal(np.concatenate([lp(theta[:7]), lp(theta[7:150]), lp(theta[150:])]), got)   # small batches too
    model = lp.model_spectrum(theta[20:24])
    ok = np.isfinite(got[20:24])
    assert np.allclose(np.median(model[ok], axis=1), np.median(p.data[1]), rtol=1e-14)
    lp.close()


def test_free_metallicity_walkers(native_lib):
    """Extension for grids beyond L2: [Fe/H] as a per-walker trailing column (7-vector).  Each walker must equal
    the fixed-[Fe/H] fit at its own metallicity (oracle on the grid interpolated there), walkers outside the
    nodes are outside the prior box, and the 7-column ensemble samples through the same stretch-move kernels."""
    import dataclasses
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.ensemble import CudaBackend, EnsembleSampler
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    p = synthetic.make_problem_3d_small(feh_free=True)
    lp = LogPosterior(p)
    assert lp.ndim == 7
    theta = synthetic.sample_prior_box(p, 96, seed=4)
    assert theta.shape == (96, 7)
    got = lp(theta)
    _compare(got, to_oracle(p).lnprob_batch(theta))
    assert got[19] == -np.inf and got[20] == -np.inf and np.isfinite(got[17]) and np.isfinite(got[18])
    # the same walker through a fixed-[Fe/H] handle: bit-identical
    k = int(np.flatnonzero(np.isfinite(got))[5])
    fixed = LogPosterior(dataclasses.replace(p, feh=float(theta[k, 6]), feh_free=False))
    assert fixed(theta[k, :6]) == got[k]
    fixed.close()
    # large batch: cell-sorted pairs / natural order agree (the cell key carries the [Fe/H] bracket)
    big = synthetic.sample_prior_box(p, 2600, seed=6, edge_cases=False)
    lp.engine.set_pairing(True); a = lp(big)
    lp.engine.set_pairing(False); b = lp(big)
    assert np.array_equal(a, b)
    # sampling in 7 dimensions
    t7 = np.concatenate([p.truth, [p.feh]])
    p0 = t7 * (1 + 0.005 * np.random.default_rng(2).standard_normal((32, 7)))
    p0[:, 2] = np.abs(p0[:, 2]); p0[:, 6] = p.feh + 0.05 * np.random.default_rng(3).standard_normal(32)
    s = EnsembleSampler(32, 7, CudaBackend(lp.engine), seed=1)
    s.initialize(p0)
    s.run_mcmc(40)
    st = s.state.cpu().numpy()
    assert np.all(np.isfinite(st[:, 7])) and s.naccepted.sum().item() > 0
    assert np.allclose(lp(st[:, :7]), st[:, 7], rtol=0, atol=0)
    lp.close()


def test_filter_count_variants(native_lib):
    """No contrast measurement at all with the three-band (2MASS-only) photometry list of mft6.py:766-767; a
    single contrast filter; four contrast + twelve photometric bands on a 2-D grid (the wide kernel variant)."""
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    base = synthetic.make_problem_small()
    wl = base.wl
    span = wl[-1] - wl[0]
    mk = lambda c, w, k: synthetic.synthetic_passband(wl[0] + c * span, w * span, 200, ripple_seed=30 + k)   # noqa: E731
    three = [mk(0.3, 0.2, 0), mk(0.55, 0.2, 1), mk(0.8, 0.25, 2)]
    cases = [
        synthetic.make_problem_small(contrast_curves=[], contrast_names=[], phot_curves=three,
                                     phot_names=['2MASS_J', '2MASS_H', '2MASS_Ks']),
        synthetic.make_problem_small(contrast_curves=[three[0]], contrast_names=['c0']),
    ]
    cc, cn, pc, pn = synthetic._many_bands(wl, 4, 12)
    cases.append(synthetic.make_problem_small(contrast_curves=cc, contrast_names=cn, phot_curves=pc, phot_names=pn))
    for prob in cases:
        lp = LogPosterior(prob)
        theta = synthetic.sample_prior_box(prob, 128, seed=3)
        _compare(lp(theta), to_oracle(prob).lnprob_batch(theta))
        parts = lp.parts(theta[:4])
        assert parts['contrast'].shape == (4, len(prob.fr[0])) and parts['phot'].shape == (4, len(prob.fr[3]))
        lp.close()


def test_c5_full_size_cell_ordered_launch(native_lib):
    """BASELINE config C5 at full size: 26x4x5 grid (416 MB > L2, so walkers are launched grouped by grid
    cell), 100 000-pixel rows, 25 000 observed pixels (one 512-thread CTA per SM), 4 + 12 bands.  Oracle on a
    sample; the launch order must not change any value (the same walkers in small batches take the natural
    order); scale invariance of the spectral chi-square."""
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    p = synthetic.make_problem_c5()
    lp = LogPosterior(p)
    assert lp.tables.n_data == 25000 and lp.tables.n_reg == 100000 and lp.tables.grid.nbytes > 400e6
    theta = synthetic.sample_prior_box(p, 2048, seed=5)
    auto = lp(theta)                                             # >= 1024 walkers: the group kernel (4-D boxes) by default
    lp.engine.set_group(False)                                   # the rest is about the per-walker kernel's launch order
    got = lp(theta)                                              # >= 1024 walkers: cell-ordered launch
    assert np.array_equal(np.isfinite(auto), np.isfinite(got))
    fin = np.isfinite(got)
    assert np.max(np.abs(auto[fin] - got[fin]) / np.abs(got[fin])) <= 1e-11
    _compare(got[:12], to_oracle(p).lnprob_batch(theta[:12]))
    pieces = np.concatenate([lp(theta[k:k + 256]) for k in range(0, 2048, 256)])   # natural order
    assert np.array_equal(pieces, got)
    a = lp.parts(theta[300:332]); th2 = theta[300:332].copy(); th2[:, 3] *= 0.8
    b = lp.parts(th2)
    ok = (a['status'] == 0) & (b['status'] == 0)
    assert ok.sum() > 8 and np.allclose(a['chi2_spec'][ok], b['chi2_spec'][ok], rtol=1e-10)
    lp.close()
    # free metallicity: the walkers spread over all five [Fe/H] planes (the HBM-bound form of the stress shape)
    import dataclasses
    pf = dataclasses.replace(p, feh_free=True)
    lpf = LogPosterior(pf)
    lpf.engine.set_group(False)
    th7 = synthetic.sample_prior_box(pf, 2048, seed=7)
    gotf = lpf(th7)
    _compare(gotf[:24], to_oracle(pf).lnprob_batch(th7[:24]))
    assert np.array_equal(np.concatenate([lpf(th7[k:k + 512]) for k in range(0, 2048, 512)]), gotf)
    lpf.close()


def test_triple_star_mode(native_lib):
    """nspec = 3: 8-parameter vectors, third gather, split contrast list
    (mft6.py:746-751), the 8-parameter prior branches (mft6.py:1329-1454)."""
    import dataclasses
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    p = synthetic.make_problem_triple_small()
    lp = LogPosterior(p)
    assert lp.ndim == 8
    theta = synthetic.sample_prior_box_triple(p, 256, seed=8)
    theta[100:] = p.truth * (1 + 0.004 * np.random.default_rng(1).standard_normal((156, 8)))
    theta[100:, 3] = np.abs(theta[100:, 3])
    orc = to_oracle(p)
    _compare(lp(theta), orc.lnprob_batch(theta))
    parts = lp.parts(theta[100:108])
    assert np.all(parts['status'] == 0)
    for k in range(8):
        ref = {}
        orc.lnprob(theta[100 + k], ref)
        assert np.allclose(parts['contrast'][k], ref['contrast'], rtol=1e-11)      # 2 x secondary, 2 x tertiary
        assert np.allclose(parts['logg'][k], ref['logg'], rtol=1e-14) and parts['logg'].shape == (8, 3)
    lp.close()
    # without a distance prior only the ratios and the sign of the parallax are boxed (mft6.py:1408)
    p2 = dataclasses.replace(p, dist_fit=False)
    lp2 = LogPosterior(p2)
    _compare(lp2(theta), to_oracle(p2).lnprob_batch(theta))
    lp2.close()


def test_broaden_rows_matches_oracle_splice(native_lib):
    from mcmc_spec_b200 import native, synthetic
    from oracle import mft6_oracle as mo
    wl = np.arange(5000.0, 9000.0, 0.2)
    rows = synthetic.make_native_grid([3000, 3100, 3200], [4.5, 5.0], wl, seed=3).reshape(6, -1).copy()
    window = [6450.0, 8400.0]
    natives = {'{}, {}'.format(k, 0.0): rows[k].copy() for k in range(6)}
    want = mo.splice_broadened_grid(natives, wl, window, 1700)
    inw = np.where((wl >= window[0]) & (wl <= window[1]))[0]
    _, taps = synthetic.gauss_taps(wl[inw][1] - wl[inw][0], np.mean(wl[inw]), 1700)
    native.broaden_rows(rows, inw[0], inw.size, taps)
    for k in range(6):
        ref = want['{}, {}'.format(k, 0.0)]
        assert np.allclose(rows[k], ref, rtol=1e-13, atol=0)
        assert np.array_equal(rows[k][:inw[0]], ref[:inw[0]])       # outside the window untouched


def test_config_errors_surface_as_exceptions(small_problem, native_lib):
    from mcmc_spec_b200 import native
    from mcmc_spec_b200.tables import build_tables
    t = build_tables(small_problem)
    t.rebin_hi = t.rebin_hi[::-1].copy()                    # not ascending
    with pytest.raises(native.NativeError):
        native.Engine(t)
