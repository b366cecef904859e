"""Host logic of the optimiser stage (fit_spec's random walk) on a cheap analytic chi-square (CPU)."""
import numpy as np

from mcmc_spec_b200 import optimizer as opt

TLIM = (3000.0, 5500.0)
BEST = np.array([4100.0, 3600.0, 0.2, 0.6, 0.5, 2.0e-3])
SCALE = np.array([100.0, 100.0, 0.05, 0.05, 0.05, 1e-4])


class Bowl:
    """chi-square = squared scaled distance to BEST; counts calls."""

    def __init__(self):
        self.calls = 0

    def start(self, theta0):
        return np.sum(((theta0 - BEST) / SCALE) ** 2, axis=1), None

    def test(self, theta, walk_index, rad_sigma):
        self.calls += len(theta)
        assert rad_sigma.shape == (len(theta), 2) and len(walk_index) == len(theta)
        return np.sum(((theta - BEST) / SCALE) ** 2, axis=1)


def test_box_is_the_reference_box():
    ok = np.array([4000.0, 3500.0, 0.1, 0.5, 0.5, 2e-3])
    assert opt._in_box(ok, TLIM)
    for k, bad in [(0, 3000.0), (0, 5500.0), (1, 2999.0), (2, -1e-9), (3, 0.049), (3, 1.51), (4, 0.05), (4, 1.0),
                   (5, 0.1), (5, 1 / 3000)]:
        v = ok.copy(); v[k] = bad
        assert not opt._in_box(v, TLIM), (k, bad)
    edge = ok.copy(); edge[2] = 0.0; edge[3] = 0.05
    assert opt._in_box(edge, TLIM)                       # Av = 0 and R1 = 0.05 are inside (mft6.py:984)


def test_fixup_only_charges_calls():
    v = np.array([2790.0, 2650.0, -0.25, 0.5, 0.02, 2e-3])
    # T: 4 rounds of +100 (2650 -> 3050), then T1 < T2? 3090 vs 3050 no; Av: 3 rounds; radius: 3 rounds (0.02 -> 0.05)
    assert opt._fixup_calls(v, TLIM, 10, 10_000) == 10 + 1 + 4 + 3 + 3
    assert opt._fixup_calls(v, TLIM, 10, 12) == 12                               # the 50*steps cap stops the loops
    assert np.array_equal(v, [2790.0, 2650.0, -0.25, 0.5, 0.02, 2e-3])           # the proposal itself is discarded
    far = np.array([4000.0, 3500.0, 0.1, 0.5, 0.5, 0.02])
    assert opt._fixup_calls(far, TLIM, 0, 10_000) == 1 + int(np.ceil(np.log(0.5) / np.log(0.99)))   # parallax shrinks 1 % a call


def test_walks_descend_and_terminate():
    rng = np.random.default_rng(3)
    theta0 = opt.draw_starts(24, *TLIM, 2e-3, 1e-4, rng)
    assert np.all(theta0[:, 1] <= theta0[:, 0]) and np.all((theta0[:, 4] > 0) & (theta0[:, 4] <= 1))
    b = Bowl()
    start_cs, _ = b.start(theta0)
    gi, chi, hist = opt.fit_spec_batch(b, theta0, TLIM, 60, rng, record=True)
    assert np.all(chi <= start_cs) and np.median(chi) < 0.05 * np.median(start_cs)
    assert np.allclose(chi, np.sum(((gi - BEST) / SCALE) ** 2, axis=1))
    assert b.calls <= 24 * 50 * 60
    for k in range(24):
        kept = [h[1] for h in hist[k]]
        assert all(x >= y for x, y in zip(kept, kept[1:]))                        # savechi never increases
        assert all(h[2] >= h[1] for h in hist[k])                                 # a kept chi is the best seen


def test_select_walkers_takes_best_third():
    params = np.arange(30).reshape(15, 2)
    cs = np.array([5, 1, 9, 3, 7, 2, 8, 4, 6, 0, 10, 11, 12, 13, 14.0])
    sel = opt.select_walkers(params, cs)
    assert sel.shape == (5, 2) and np.array_equal(sel[:, 0] // 2, [9, 1, 5, 3, 7])
