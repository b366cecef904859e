"""Sampler host logic on CPUs: stretch-move statistics and walker sharding under gloo (world size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mcmc_spec_b200.ensemble import EnsembleSampler
import stretch_reference as sr

NDIM = 4
MU = np.array([1.0, -2.0, 0.5, 3.0])
SIG = np.array([0.5, 2.0, 1.0, 0.1])


def gauss_lnprob(theta):
    out = -0.5 * np.sum(((theta - MU) / SIG) ** 2, axis=1)
    out[theta[:, 0] > 2.4] = -np.inf                     # a hard prior edge: -inf proposals must be rejected
    return out


def _p0(nw, seed=3):
    return MU + SIG * np.random.default_rng(seed).standard_normal((nw, NDIM)) * 0.5


def test_counter_uniform_is_uniform_and_reproducible():
    u = sr.counter_uniform(7, 3, 1, 0, np.arange(200000))
    assert 0 < u.min() and u.max() < 1
    assert abs(u.mean() - 0.5) < 3e-3 and abs(u.var() - 1 / 12) < 2e-3
    assert np.array_equal(u[:10], sr.counter_uniform(7, 3, 1, 0, np.arange(10)))
    assert not np.array_equal(u[:10], sr.counter_uniform(7, 4, 1, 0, np.arange(10)))
    # z = ((a-1)u+1)^2/a has density ~ 1/sqrt(z) on [1/a, a]: E[z] = (a^2 + a + 1) / (3a)
    z = ((2 - 1) * u + 1) ** 2 / 2
    assert abs(z.mean() - 7 / 6) < 3e-3


def test_chain_recovers_gaussian_moments():
    s = EnsembleSampler(64, NDIM, sr.NumpyBackend(gauss_lnprob, NDIM), seed=11)
    s.initialize(_p0(64))
    s.run_mcmc(300)
    samples = []
    for _ in range(700):
        s.step()
        samples.append(s.coords.numpy().copy())
    x = np.concatenate(samples)
    assert np.all(x[:, 0] <= 2.4)                                        # the -inf region is never entered
    assert np.allclose(x[:, 1:].mean(0), MU[1:], atol=4 * SIG[1:] / np.sqrt(700))      # ~700 x 64 / tau samples
    assert np.allclose(x[:, 1:].std(0), SIG[1:], rtol=0.08)
    af = s.acceptance_fraction
    assert 0.3 < af.mean() < 0.8
    assert np.all(np.isfinite(s.log_prob.numpy()))


def test_fixed_split_and_state_consistency():
    s = EnsembleSampler(32, NDIM, sr.NumpyBackend(gauss_lnprob, NDIM), seed=2, randomize_split=False)
    s.initialize(_p0(32))
    for _ in range(20):
        s.step()
    assert np.allclose(s.log_prob.numpy(), gauss_lnprob(s.coords.numpy().copy()), rtol=0, atol=0)
    with pytest.raises(ValueError):
        EnsembleSampler(31, NDIM, sr.NumpyBackend(gauss_lnprob, NDIM))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nw, nsteps, out, by_group=False):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    backend = (sr.GroupShardedNumpyBackend if by_group else sr.NumpyBackend)(gauss_lnprob, NDIM)
    s = EnsembleSampler(nw, NDIM, backend, seed=5, group=None, rank=rank, world_size=world)
    if by_group:
        s.shard_by_group = 'always'
    s.initialize(_p0(nw))
    s.run_mcmc(nsteps)
    if rank == 0:
        np.save(out, np.concatenate([s.state.numpy(), s.naccepted.numpy()[:, None].astype(np.float64)], axis=1))
    else:
        # every rank holds the same replicated state
        ref = [None]
    full = [torch.empty_like(s.state) for _ in range(world)]
    dist.all_gather(full, s.state)
    assert all(torch.equal(f, full[0]) for f in full)
    assert bool(s._by_group) == bool(by_group)
    dist.destroy_process_group()


def test_two_rank_gloo_chain_is_bit_identical_to_single_process(tmp_path):
    nw, nsteps = 48, 25
    single = EnsembleSampler(nw, NDIM, sr.NumpyBackend(gauss_lnprob, NDIM), seed=5)
    single.initialize(_p0(nw))
    single.run_mcmc(nsteps)
    out = str(tmp_path / 'state.npy')
    mp.spawn(_worker, args=(2, _free_port(), nw, nsteps, out), nprocs=2, join=True)
    got = np.load(out)
    assert np.array_equal(got[:, :NDIM + 1], single.state.numpy())        # independent of the number of ranks
    assert np.array_equal(got[:, NDIM + 1], single.naccepted.numpy())


def test_two_rank_gloo_chain_sharded_by_groups_is_the_same_chain(tmp_path):
    """The other way N ranks share a half step (ensemble.py): every rank holds the ensemble,
    proposes for the whole half, evaluates every N-th group, the log-probs are summed by one all-reduce and every
    rank accepts in place.  Same chain as one process, same state on both ranks."""
    nw, nsteps = 48, 25
    single = EnsembleSampler(nw, NDIM, sr.NumpyBackend(gauss_lnprob, NDIM), seed=5)
    single.initialize(_p0(nw))
    single.run_mcmc(nsteps)
    out = str(tmp_path / 'state_g.npy')
    mp.spawn(_worker, args=(2, _free_port(), nw, nsteps, out, True), nprocs=2, join=True)
    got = np.load(out)
    assert np.array_equal(got[:, :NDIM + 1], single.state.numpy())
    assert np.array_equal(got[:, NDIM + 1], single.naccepted.numpy())


def test_emcee_style_calls_of_the_reference(tmp_path):
    """The commented sampling block of the reference (mft6.py:1491-1527) written against this sampler: sample()
    generator with .coords, get_last_sample(), reset(), get_autocorr_time(quiet=True), acceptance_fraction and the
    [nwalkers, nsteps, ndim] chain."""
    nw = 24
    sampler = EnsembleSampler(nw, NDIM, sr.NumpyBackend(gauss_lnprob, NDIM), seed=4)
    n_dumps = 0
    for n, s in enumerate(sampler.sample(_p0(nw), iterations=30)):
        if n % 10 == 0:
            with open(tmp_path / 'x_{}_burnin.txt'.format(n), 'ab') as f:
                f.write(b"\\n"); np.savetxt(f, s.coords)
            n_dumps += 1
    assert n_dumps == 3
    state = sampler.get_last_sample()
    before = state.coords.clone()
    sampler.reset()
    assert sampler.get_chain().shape == (0, nw, NDIM) and np.all(sampler.acceptance_fraction == 0)
    assert torch.equal(sampler.coords, before)
    for n, s in enumerate(sampler.sample(state, iterations=200)):
        pass
    acl = sampler.get_autocorr_time(quiet=True)
    assert acl.shape == (NDIM,) and np.all(np.isfinite(acl)) and np.all(acl > 0.5)
    assert 0.2 < np.mean(sampler.acceptance_fraction) < 0.9
    chain = sampler.stored_chain
    assert chain.shape == (nw, 200, NDIM)
    samples = chain[:, :, :].reshape((-1, NDIM))                      # mft6.py:1527
    assert np.array_equal(samples[-1], sampler.coords.numpy()[-1])
    # same chain as run_mcmc with the same seed and start
    ref = EnsembleSampler(nw, NDIM, sr.NumpyBackend(gauss_lnprob, NDIM), seed=4)
    ref.initialize(_p0(nw)); ref.run_mcmc(30)
    assert np.array_equal(ref.coords.numpy(), before.numpy())
