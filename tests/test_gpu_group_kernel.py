"""The group kernel (lnprob_group_kernel: cell-mates share one TMA-staged stream of their rows' bracket pairs, lane = walker)
against the oracle and against the one-CTA-per-walker kernel, on the cases where its own code paths differ: exact
median over odd / even / tiny / ragged sizes and many equal keys (per-lane histogram + key lists), partial groups,
walkers outside the prior box, both kernel variants (8 / 16 walker warps), evaluation sharded by groups over ranks,
intermediate outputs, determinism.

Bar: identical -inf sets, 1e-9 relative against the oracle (tolerance of north_star), 1e-11 between the two device
paths (they differ in the summation order of the band integrals and the moments).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    f = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), f)
    assert not np.any(np.isnan(a))
    return float(np.max(np.abs(a[f] - b[f]) / np.abs(b[f]))) if f.any() else 0.0


def _both(lp, theta):
    lp.engine.set_group(True)
    g = lp(theta)
    g2 = lp(theta)
    lp.engine.set_group(False)
    s = lp(theta)
    assert np.array_equal(g, g2)                                 # deterministic (grouping is not: atomics order the cells)
    return g, s


@pytest.mark.parametrize('n_data', [1, 2, 7, 63, 64, 255, 256, 300, 301, 1025])
def test_median_sizes_through_the_group_kernel(native_lib, n_data):
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    p = synthetic.make_problem_small(n_data=n_data, seed=40 + n_data)
    lp = LogPosterior(p)
    assert lp.engine.group_supported
    theta = synthetic.sample_prior_box(p, 200, seed=n_data)
    theta[100:] = np.asarray(p.truth) * (1 + 0.003 * np.random.default_rng(n_data).standard_normal((100, 6)))   # one dense cell
    theta[100:, 2] = np.abs(theta[100:, 2])
    g, s = _both(lp, theta)
    lp.close()
    assert _rel(g, to_oracle(p).lnprob_batch(theta)) <= 1e-9
    assert _rel(g, s) <= 1e-11


def test_equal_keys_and_huge_av_through_the_group_kernel(native_lib):
    """Flat rows: every model value equal (Av = 0) or a smooth ramp (Av > 0); Av far beyond the polynomial's range
    takes the exp2 path (those walkers only)."""
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    p = synthetic.make_problem_small(n_data=200, seed=77)
    p.grid[...] = 1.0e6
    lp = LogPosterior(p)
    theta = synthetic.sample_prior_box(p, 96, seed=5)
    theta[:16, 2] = 0.0
    theta[16:24, 2] = 9.5                                         # |Av x| >> 0.045
    g, s = _both(lp, theta)
    lp.close()
    assert _rel(g, to_oracle(p).lnprob_batch(theta)) <= 1e-9
    assert _rel(g, s) <= 1e-11


def test_parts_and_model_through_the_group_kernel(small_problem, small_oracle, native_lib):
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    lp = LogPosterior(small_problem)
    lp.engine.set_group(True)
    theta = synthetic.sample_prior_box(small_problem, 80, seed=9)
    parts = lp.parts(theta)
    model = lp.model_spectrum(theta[16:24])
    for k in range(80):
        ref = {}
        v = small_oracle.lnprob(theta[k], ref)
        if not np.isfinite(v):
            assert parts['lnprob'][k] == -np.inf and parts['status'][k] != 0
            continue
        assert parts['lnprior'][k] == pytest.approx(ref['lp'], rel=1e-12)
        assert parts['chi2_spec'][k] == pytest.approx(ref['chi2_spec'], rel=1e-11)
        assert np.allclose(parts['contrast'][k], ref['contrast'], rtol=1e-12)
        assert np.allclose(parts['phot'][k], ref['phot'], rtol=1e-13)
        assert np.allclose(parts['logg'][k], ref['logg'], rtol=1e-14)
        if 16 <= k < 24:
            assert np.allclose(model[k - 16], ref['model'], rtol=1e-13)
    lp.close()


def test_both_kernel_variants_give_the_same_bits(native_lib):
    """Few groups -> sixteen walker warps per group on one CTA per SM, many groups -> eight on two CTAs per SM; a
    walker's value must not depend on the variant, the batch it arrives in or its companions in the group (the
    sharded chain then walks the same path on any number of GPUs)."""
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    p = synthetic.make_problem_small()
    lp = LogPosterior(p)
    lp.engine.set_group(True)
    rng = np.random.default_rng(3)
    theta = np.asarray(p.truth) * (1 + 0.01 * rng.standard_normal((16384, 6)))
    theta[:, 2] = np.abs(theta[:, 2])
    big = lp(theta)
    big = lp(theta)                                               # second call: the variant follows the group count seen
    small = np.concatenate([lp(theta[k:k + 1024]) for k in range(0, 16384, 1024)])
    small = np.concatenate([lp(theta[k:k + 1024]) for k in range(0, 16384, 1024)])
    perm = rng.permutation(16384)
    shuffled = lp(theta[perm])
    lp.close()
    assert np.array_equal(big, small)
    assert np.array_equal(big[perm], shuffled)


def test_group_shards_sum_to_the_whole_half_step(native_lib):
    """ms_set_group_shard: N ranks that hold the same ensemble each evaluate every N-th group of the whole half; the
    sum of their log-probs (each walker non-zero on exactly one rank) is bit for bit what one rank computes alone,
    the proposals are the same on every rank, and the in-place acceptance gives the single-rank ensemble."""
    import torch
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.ensemble import CudaBackend
    from mcmc_spec_b200.posterior import LogPosterior
    p = synthetic.make_problem_small()
    lp = LogPosterior(p)
    lp.engine.set_group(True)
    be = CudaBackend(lp.engine)
    rng = np.random.default_rng(11)
    nw = 4096
    theta = np.asarray(p.truth) * (1 + 0.02 * rng.standard_normal((nw, 6)))
    theta[:, 2] = np.abs(theta[:, 2])
    state = torch.tensor(np.concatenate([theta, lp(theta)[:, None]], axis=1), device='cuda')
    s_idx = torch.arange(0, nw, 2, device='cuda'); c_idx = torch.arange(1, nw, 2, device='cuda')
    n_s = s_idx.shape[0]
    f64 = dict(dtype=torch.float64, device='cuda')

    def evaluate(rank, world):
        q = torch.empty((n_s, 6), **f64); lf = torch.empty(n_s, **f64); lnp = torch.full((n_s,), 123.0, **f64)
        be.set_shard(rank, world)
        be.evaluate(state, s_idx, c_idx, 2.0, 5, 3, 0, 0, q, lf, lnp)
        return q, lf, lnp

    q1, lf1, whole = evaluate(0, 1)
    parts = [evaluate(r, 3) for r in range(3)]
    be.set_shard(0, 1)
    for q, lf, _ in parts:
        assert torch.equal(q, q1) and torch.equal(lf, lf1)
    nonzero = torch.stack([(l != 0) for _, _, l in parts]).sum(dim=0)
    assert int(nonzero.max()) == 1 and int((nonzero == 0).sum()) == 0          # every walker on exactly one rank (no lnprob is 0 here)
    total = parts[0][2] + parts[1][2] + parts[2][2]
    assert torch.equal(total, whole)
    for _, _, l in parts:
        assert 0.2 < float((l != 0).double().mean()) < 0.5                     # roughly a third of the groups each
    # acceptance in place == the one-call half step
    st_a = state.clone(); st_b = state.clone()
    na = torch.zeros(nw, dtype=torch.int32, device='cuda'); nb = torch.zeros(nw, dtype=torch.int32, device='cuda')
    be.accept_inplace(st_a, na, s_idx, q1, total, lf1, 5, 3, 0, 0)
    q = torch.empty((n_s, 6), **f64); lf = torch.empty(n_s, **f64); lnp = torch.empty(n_s, **f64)
    be.halfstep(st_b, nb, s_idx, c_idx, 2.0, 5, 3, 0, 0, q, lf, lnp)
    assert torch.equal(st_a, st_b) and torch.equal(na, nb) and int(na.sum()) > 0
    lp.close()


@pytest.mark.parametrize('feh_free', [False, True])
def test_feh_axis_through_the_group_kernel(native_lib, feh_free):
    """Grids with an [Fe/H] axis (BASELINE config C5): eight bracketing rows per star arrive as one 4-D tensor box;
    fixed [Fe/H] (on and off a node) and [Fe/H] as a walker parameter (the cell key carries its bracket)."""
    from mcmc_spec_b200 import synthetic
    from mcmc_spec_b200.posterior import LogPosterior
    from oracle_adapter import to_oracle
    for feh in ((0.2,) if feh_free else (0.2, 0.0, -0.5)):
        p = synthetic.make_problem_3d_small(feh_free=True) if feh_free else synthetic.make_problem_3d_small(feh=feh)
        lp = LogPosterior(p)
        assert lp.engine.group_supported
        theta = synthetic.sample_prior_box(p, 320, seed=17)
        t = np.asarray(p.truth, dtype=float)
        dense = t[None, :] * (1 + 0.004 * np.random.default_rng(5).standard_normal((160, t.size)))     # one dense cell
        dense[:, 2] = np.abs(dense[:, 2])
        theta[160:, :t.size] = dense
        g, s = _both(lp, theta)
        lp.close()
        assert _rel(g, to_oracle(p).lnprob_batch(theta)) <= 1e-9
        assert _rel(g, s) <= 1e-11
