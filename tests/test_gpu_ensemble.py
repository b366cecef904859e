"""GPU: stretch-move kernels against their numpy mirror, and a fixed-seed chain against the oracle chain."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def small_lp(small_problem, native_lib):
    from mcmc_spec_b200.posterior import LogPosterior
    lp = LogPosterior(small_problem)
    yield lp
    lp.close()


def test_sampler_kernels_match_numpy(small_lp):
    import torch
    import stretch_reference as sr
    from mcmc_spec_b200.ensemble import CudaBackend
    be = CudaBackend(small_lp.engine)
    rng = np.random.default_rng(0)
    nw, nd = 1400, 6
    state = rng.normal(size=(nw, nd + 1))
    keys = be.split_keys(42, 17, nw).cpu().numpy()
    assert np.array_equal(keys, sr.split_keys(42, 17, nw)) and keys.min() >= 0 and len(set(keys)) == nw
    perm = np.argsort(keys, kind='stable')
    s_idx, c_idx = perm[0::2][100:600].copy(), perm[1::2].copy()
    dev = lambda a: torch.tensor(a, device='cuda')                         # noqa: E731
    st = dev(state)
    q = torch.empty((500, nd), dtype=torch.float64, device='cuda'); lf = torch.empty(500, dtype=torch.float64, device='cuda')
    be.propose(st, dev(s_idx), dev(c_idx), 2.0, 42, 17, 1, 100, q, lf)
    q_ref, lf_ref = sr.propose(state[s_idx, :nd], state[c_idx, :nd], 2.0, 42, 17, 1, 100)
    assert np.allclose(q.cpu().numpy(), q_ref, rtol=1e-15, atol=1e-15)       # same partner, same z
    assert np.allclose(lf.cpu().numpy(), lf_ref, rtol=1e-13, atol=1e-15)
    lnq = rng.normal(size=500); lnq[::9] = -np.inf
    packed = torch.empty((500, nd + 2), dtype=torch.float64, device='cuda')
    be.accept(st, dev(s_idx), q, dev(lnq), lf, 42, 17, 1, 100, packed)
    mask = sr.accept_mask(state[s_idx, nd], lnq, lf_ref, 42, 17, 1, 100)
    pk = packed.cpu().numpy()
    assert np.array_equal(pk[:, nd + 1].astype(bool), mask) and not mask[::9].any() and mask.any()
    assert np.array_equal(pk[mask, :nd], q.cpu().numpy()[mask]) and np.array_equal(pk[~mask, :nd + 1], state[s_idx][~mask])
    assert np.array_equal(pk[:, nd], np.where(mask, lnq, state[s_idx, nd]))
    nacc = torch.zeros(nw, dtype=torch.int32, device='cuda')
    be.scatter(st, nacc, dev(s_idx), packed)
    want = state.copy(); want[s_idx] = pk[:, :nd + 1]
    assert np.array_equal(st.cpu().numpy(), want)
    assert np.array_equal(nacc.cpu().numpy()[s_idx], mask.astype(np.int32)) and nacc.sum().item() == mask.sum()


def test_fixed_seed_chain_matches_oracle_chain(small_problem, small_oracle, small_lp):
    """Same z / partner / u streams fed to the CUDA path and to the oracle: the chains coincide."""
    import stretch_reference as sr
    from mcmc_spec_b200.ensemble import CudaBackend, EnsembleSampler
    nw, nsteps = 32, 120
    t = small_problem.truth
    p0 = t * (1 + 0.01 * np.random.default_rng(1).standard_normal((nw, 6)))
    p0[:, 2] = np.abs(p0[:, 2])
    gpu = EnsembleSampler(nw, 6, CudaBackend(small_lp.engine), seed=9)
    cpu = EnsembleSampler(nw, 6, sr.NumpyBackend(small_oracle.lnprob_batch, 6), seed=9)
    gpu.initialize(p0); cpu.initialize(p0)
    gpu.run_mcmc(nsteps); cpu.run_mcmc(nsteps)
    g, c = gpu.state.cpu().numpy(), cpu.state.numpy()
    assert np.array_equal(gpu.naccepted.cpu().numpy(), cpu.naccepted.numpy())      # identical decisions
    assert np.allclose(g[:, :6], c[:, :6], rtol=1e-12)
    assert np.allclose(g[:, 6], c[:, 6], rtol=1e-9)
    assert gpu.naccepted.sum().item() > 0
    # ... and so do the posterior summaries the reference reports (median, 16th/84th percentiles)
    gs, cs = [], []
    for _ in range(30):
        gpu.step(); cpu.step()
        gs.append(gpu.coords.cpu().numpy().copy()); cs.append(cpu.coords.numpy().copy())
    gq = np.percentile(np.concatenate(gs), [16, 50, 84], axis=0)
    cq = np.percentile(np.concatenate(cs), [16, 50, 84], axis=0)
    assert np.allclose(gq, cq, rtol=1e-10)


def test_independent_chains_agree_within_monte_carlo_error(small_problem, small_oracle, small_lp):
    """Independently seeded chains -- CUDA posterior + device sampler vs numpy oracle + numpy sampler -- give
    the same posterior medians and 16/84 percentiles within the stated Monte-Carlo tolerance:
    |delta| <= 4 sigma_MC, sigma_MC = 1.25 * std * sqrt(tau / n_samples) per chain, added in quadrature
    (1.25: a percentile's standard error relative to the mean's, normal approximation)."""
    import stretch_reference as sr
    from mcmc_spec_b200 import fit
    from mcmc_spec_b200.ensemble import CudaBackend, EnsembleSampler
    nw = 24
    t = small_problem.truth
    p0 = t * (1 + 0.005 * np.random.default_rng(8).standard_normal((nw, 6)))
    p0[:, 2] = np.abs(p0[:, 2])

    def run(sampler, start, burn, steps):
        sampler.initialize(start)
        sampler.run_mcmc(burn)
        out = []
        for _ in range(steps):
            sampler.step()
            c = sampler.coords
            out.append((c.cpu().numpy() if c.is_cuda else c.numpy()).copy())
        return np.array(out)                                            # [steps, nw, ndim]

    g = run(EnsembleSampler(nw, 6, CudaBackend(small_lp.engine), seed=101), p0, 1000, 6000)
    # the (slow) oracle chain starts from an equilibrated ensemble -- the device chain's last state -- with its
    # own random streams; tau is taken from the long chain (same dynamics), a 300-step chain cannot estimate it
    c = run(EnsembleSampler(nw, 6, sr.NumpyBackend(small_oracle.lnprob_batch, 6), seed=202), g[-1], 20, 300)
    tau = np.maximum(fit.integrated_autocorr_time(g), 1.0)
    gq = np.percentile(g.reshape(-1, 6), [16, 50, 84], axis=0)
    cq = np.percentile(c.reshape(-1, 6), [16, 50, 84], axis=0)
    std = g.reshape(-1, 6).std(axis=0)
    sig = 1.25 * std * np.sqrt(tau / (g.shape[0] * nw) + tau / (c.shape[0] * nw))
    assert np.all(np.abs(gq - cq) <= 4.0 * sig), (np.abs(gq - cq) / sig, tau)
    # the comparison has teeth: the tolerance is a fraction of the posterior width
    assert np.all(4.0 * sig < 1.5 * std)


@pytest.mark.parametrize('nw,randomize', [(50, True), (32, False), (2048, True), (6, True)])
def test_fused_iteration_is_the_same_chain(small_problem, small_lp, nw, randomize):
    """ms_stretch_iteration (one call per iteration: in-kernel split sort + both half steps) walks exactly the
    chain of the per-kernel path (torch argsort + five launches per half step): three launches per half step --
    proposal folded into the prep kernel, streaming kernel, in-place accept."""
    from mcmc_spec_b200.ensemble import CudaBackend, EnsembleSampler
    t = small_problem.truth
    p0 = t * (1 + 0.01 * np.random.default_rng(nw).standard_normal((nw, 6)))
    p0[:, 2] = np.abs(p0[:, 2])
    a = EnsembleSampler(nw, 6, CudaBackend(small_lp.engine), seed=5, randomize_split=randomize)
    b = EnsembleSampler(nw, 6, CudaBackend(small_lp.engine), seed=5, randomize_split=randomize)
    b.fused = False
    a.initialize(p0); b.initialize(p0)
    n0 = small_lp.engine.launch_count
    a.run_mcmc(12)
    n1 = small_lp.engine.launch_count
    b.run_mcmc(12)
    assert np.array_equal(a.state.cpu().numpy(), b.state.cpu().numpy())
    assert np.array_equal(a.naccepted.cpu().numpy(), b.naccepted.cpu().numpy())
    assert a.naccepted.sum().item() > 0 and a.iteration == b.iteration == 12
    # split + 2 x (propose+prep, stream, accept in place); half steps of >= 1024 walkers go through the group kernel,
    # which adds the cell sort and the group table (scan, scatter, table)
    per_half = 3 + (3 if nw // 2 >= 1024 and small_lp.engine.group_supported else 0)
    assert n1 - n0 == 12 * (1 + 2 * per_half)


def test_emcee_style_calls_on_the_device_sampler(small_problem, small_lp):
    """sample() / get_last_sample() / reset() / get_autocorr_time() / stored_chain on the CUDA backend."""
    from mcmc_spec_b200.ensemble import CudaBackend, EnsembleSampler
    t = small_problem.truth
    p0 = t * (1 + 0.01 * np.random.default_rng(3).standard_normal((50, 6)))
    p0[:, 2] = np.abs(p0[:, 2])
    sampler = EnsembleSampler(50, 6, CudaBackend(small_lp.engine), seed=12)
    for n, s in enumerate(sampler.sample(p0, iterations=40)):
        assert s.coords.shape == (50, 6)
    state = sampler.get_last_sample()
    sampler.reset()
    for n, s in enumerate(sampler.sample(state, iterations=300)):
        pass
    acl = sampler.get_autocorr_time(quiet=True)
    chain = sampler.stored_chain
    assert chain.shape == (50, 300, 6) and acl.shape == (6,) and np.all(np.isfinite(acl))
    assert 0.1 < sampler.acceptance_fraction.mean() < 0.9 and sampler.iteration == 340
    assert np.allclose(small_lp(chain[:, -1]), sampler.log_prob.cpu().numpy(), rtol=0, atol=0)
