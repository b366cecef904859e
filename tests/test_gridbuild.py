"""Grid builder (spec_interpolator mirror): host logic on CPUs, kernels on the GPU."""
import os

import numpy as np
import pytest

from mcmc_spec_b200 import gridbuild


def _write_nodes(d, temps, lgs, wl_lo=2900.0, wl_hi=6100.0, seed=0):
    """Small BT-Settl-style directory: one two-column text file per node, irregular native sampling."""
    rng = np.random.default_rng(seed)
    truth = {}
    for t in temps:
        for g in lgs:
            n = int(rng.integers(5000, 6000))
            w = np.sort(rng.uniform(wl_lo, wl_hi, n))
            w[0], w[-1] = wl_lo, wl_hi
            f = 1e5 * (t / 4000.0) ** 4 * (1 + 0.3 * np.sin(w / 37.0 + g) + 0.05 * rng.standard_normal(n))
            name = 'lte{}-{}-0.0a+0.0.BT-Settl.spec.7.txt'.format(str(int(t / 1e2)).zfill(3), str(g))
            with open(os.path.join(d, name), 'w') as fh:
                for a, b in zip(w, f):
                    fh.write('{} {} 0.0\n'.format(repr(float(a)), repr(float(b))))
            truth[(t, g)] = (w, f)
    return truth


def test_discover_and_read_nodes(tmp_path):
    temps, lgs = [3000, 3100, 3200, 4000], [4.5, 5.0]
    truth = _write_nodes(str(tmp_path), temps, lgs)
    t, l = gridbuild.discover_nodes(str(tmp_path), [3000, 3200], [4, 5.5])
    assert sorted(t) == [3000, 3100, 3200] and sorted(l) == [4.5, 5.0]          # 4000 K is outside trange
    path = gridbuild.node_file(str(tmp_path), 3100, 5.0)
    assert path.endswith('lte031-5.0-0.0a+0.0.BT-Settl.spec.7.txt')
    w, f = gridbuild.read_node(path, [3000, 6000])
    tw, tf = truth[(3100, 5.0)]
    assert np.array_equal(w, tw) and np.array_equal(f, tf)                      # +-100 A margin keeps everything
    w2, _ = gridbuild.read_node(path, [3500, 5000])
    assert w2.min() >= 3400 and w2.max() <= 5100 and w2.size < w.size
    with pytest.raises(FileNotFoundError):
        gridbuild.node_file(str(tmp_path), 5000, 5.0)


def test_gaussian_taps_match_the_oracle_kernel():
    from oracle import thirdparty as tp
    wl = np.arange(6450.0, 8400.0, 0.2)
    taps = gridbuild.gaussian_taps(wl, 1700)
    assert taps.size == 93 and abs(taps.sum() - 1.0) < 1e-15
    sigma = np.mean(wl) / 1700.0 / (2.0 * np.sqrt(2.0 * np.log(2.0)))
    assert np.allclose(taps, tp.gauss_kernel_taps(wl[1] - wl[0], sigma, 5)[1], rtol=1e-12)
    with pytest.raises(ValueError):
        gridbuild.gaussian_taps(np.array([1.0, 2.0, 4.0]), 1000)


@pytest.mark.gpu
def test_spec_interpolator_on_gpu_matches_numpy_and_oracle(tmp_path, native_lib):
    from scipy.interpolate import interp1d
    from oracle import mft6_oracle as mo
    temps, lgs = [3000, 3100, 3200], [4.5, 5.0]
    truth = _write_nodes(str(tmp_path), temps, lgs)
    window, res = [4000.0, 5000.0], 1700
    grid, wl, t_out, g_out = gridbuild.spec_interpolator(str(tmp_path), window, [3000, 3200], [4, 5.5], [3000, 6000], res)
    assert t_out == temps and g_out == lgs and grid.shape == (2, 3, wl.size)
    assert np.array_equal(wl, np.arange(3000, 6000, 0.2))
    inwin = (wl >= 4000.0) & (wl <= 5000.0)
    for ig, g in enumerate(lgs):
        for it, t in enumerate(temps):
            w, f = truth[(t, g)]
            native = interp1d(w, f)(wl)                                         # mft6.py:369-371
            assert np.array_equal(grid[ig, it][~inwin], native[~inwin])        # resampling is bit-for-bit np.interp
            want = mo.splice_broadened_grid({'k': native}, wl, window, res)['k']   # mft6.py:373-376
            assert np.max(np.abs(grid[ig, it] - want) / np.abs(want)) < 1e-13
    # interp1d's bounds_error: the common grid must lie inside every node's native range
    with pytest.raises(ValueError):
        gridbuild.spec_interpolator(str(tmp_path), window, [3000, 3200], [4, 5.5], [2000, 6000], res)
