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


def test_propose_and_accept_kernels_match_numpy(small_lp):
    import torch
    import stretch_reference as sr
    from mcmc_spec_b200.ensemble import CudaBackend
    be = CudaBackend(small_lp.engine)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(777, 6)); c = rng.normal(size=(500, 6))
    q, lf = be.propose(torch.tensor(x, device='cuda'), torch.tensor(c, device='cuda'), 2.0, 42, 17, 1, 1000)
    q_ref, lf_ref = sr.propose(x, c, 2.0, 42, 17, 1, 1000)
    assert np.allclose(q.cpu().numpy(), q_ref, rtol=1e-15, atol=1e-15)       # same partner, same z
    assert np.allclose(lf.cpu().numpy(), lf_ref, rtol=1e-13, atol=1e-15)
    lnx = rng.normal(size=777); lnq = rng.normal(size=777); lnq[::9] = -np.inf
    xt, lt = torch.tensor(x, device='cuda'), torch.tensor(lnx, device='cuda')
    acc = torch.zeros(777, dtype=torch.int32, device='cuda')
    be.accept(xt, lt, q, torch.tensor(lnq, device='cuda'), lf, 42, 17, 1, 1000, acc)
    mask = sr.accept_mask(lnx, lnq, lf_ref, 42, 17, 1, 1000)
    assert np.array_equal(acc.cpu().numpy().astype(bool), mask) and not mask[::9].any()
    assert np.array_equal(xt.cpu().numpy()[mask], q.cpu().numpy()[mask]) and np.array_equal(xt.cpu().numpy()[~mask], x[~mask])
    assert np.array_equal(lt.cpu().numpy(), np.where(mask, lnq, lnx))


def test_fixed_seed_chain_matches_oracle_chain(small_problem, small_oracle, small_lp):
    """Same z / partner / u streams fed to the CUDA path and to the oracle: the chains coincide."""
    import stretch_reference as sr
    from mcmc_spec_b200.ensemble import CudaBackend, EnsembleSampler
    nw, nsteps = 32, 25
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
