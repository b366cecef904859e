"""The oracle reproduces the committed golden vectors (CPU).  See tests/golden/make_fixtures.py."""
import os

import numpy as np
import pytest

import golden_io
from oracle_adapter import to_oracle

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.mark.parametrize('name,n_data', [('c2_koi2298', 1945), ('c1_synth_3850_3425', 2183)])
def test_oracle_matches_golden(name, n_data):
    path = os.path.join(GOLD, name + '.npz')
    fx = np.load(path)
    prob = golden_io.problem_from_fixture(path)
    assert prob.n_data == n_data                      # N after mask + cut
    assert prob.grid.shape == (4, 26, 135000)
    orc = to_oracle(prob)
    idx = list(range(0, 64, 3))                       # a third of the vectors keeps the CPU suite short
    got = orc.lnprob_batch(fx['theta'][idx])
    want = fx['oracle_lnprob'][idx]
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    f = np.isfinite(want)
    assert np.allclose(got[f], want[f], rtol=1e-12, atol=0)


def test_oracle_matches_c3_golden():
    """Config C3 fixture (12 synth spectra, one grid): spot-check two of the fits on the CPU."""
    import dataclasses
    fx = np.load(os.path.join(GOLD, 'c3_synth_all.npz'))
    assert len(fx['names']) == 12 and set(fx['lens']) == {2183}
    base = golden_io.problem_from_fixture(os.path.join(GOLD, 'c1_synth_3850_3425.npz'))
    off = np.concatenate([[0], np.cumsum(fx['lens'])])
    for k in (0, 7):
        sl = slice(off[k], off[k + 1])
        prob = dataclasses.replace(base, data=[fx['wl'][sl], fx['flux'][sl]], err=fx['err'][sl])
        got = to_oracle(prob).lnprob_batch(fx['theta'][k][:3])
        want = fx['oracle_lnprob'][k][:3]
        f = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), f) and np.allclose(got[f], want[f], rtol=1e-12, atol=0)


@pytest.mark.parametrize('name', ['c1', 'c2', 'c3', 'c4', 'c5'])
def test_oracle_matches_10k_fixture(name):
    """tests/golden/c*_10k.npz (make_theta10k.py): the oracle reproduces a spot sample of each file on the CPU,
    so a stale fixture or a changed oracle is caught without a GPU."""
    import sys
    sys.path.insert(0, GOLD)
    import make_theta10k
    fx = np.load(os.path.join(GOLD, name + '_10k.npz'))
    assert len(fx['theta']) >= 10000 and fx['theta'].shape[1] == 6
    assert 9000 < np.isfinite(fx['oracle_lnprob']).sum() < len(fx['theta'])       # edge cases give -inf
    idx = [0, 8, 9, 4001, 8500, 9999] if name != 'c3' else [0, 8, 700]
    k = int(fx['problem'][idx[0]]) if name == 'c3' else 0
    orc = to_oracle(make_theta10k.problem_for(name, k))
    got = orc.lnprob_batch(fx['theta'][idx])
    want = fx['oracle_lnprob'][idx]
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    f = np.isfinite(want)
    assert np.allclose(got[f], want[f], rtol=1e-12, atol=0)
