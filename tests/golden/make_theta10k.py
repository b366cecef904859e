"""Generate tests/golden/c{1..5}_10k.npz: 10^4 parameter vectors per BASELINE config with the oracle's
log-posterior (">= 10^4 random theta per config incl. edge cases").

    python tests/golden/make_theta10k.py [c1 c2 c3 c4 c5]      # BUILD container, all host cores, minutes

Each file holds `theta` [10000, 6] and `oracle_lnprob` [10000] (C3: also `problem` [10000], the index of the
Data/synth_spec_* spectrum the vector is evaluated against).  The problems themselves are rebuilt at test time:
C1/C2/C3 from the committed fixtures of make_fixtures.py (reference data files), C4/C5 from the seeded
synthetic generators (mcmc_spec_b200.synthetic, seed 2298).  8000 vectors are uniform in the prior box of
SURVEY.md section 8(d) with the forced edge cases of synthetic.sample_prior_box (on-node / on-edge Teff,
Av = 0, Av < 0, out-of-box radii and parallaxes, find_nearest ties); 2000 are a 2 % cloud around a plausible
solution, where the likelihood rather than the prior decides the value.
"""
import dataclasses
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

N_BOX, N_CLOUD = 8000, 2000


def problem_for(name, k=0):
    from mcmc_spec_b200 import synthetic
    from golden_io import problem_from_fixture
    if name == 'c1':
        return problem_from_fixture(os.path.join(HERE, 'c1_synth_3850_3425.npz'))
    if name == 'c2':
        return problem_from_fixture(os.path.join(HERE, 'c2_koi2298.npz'))
    if name == 'c3':
        base = problem_from_fixture(os.path.join(HERE, 'c1_synth_3850_3425.npz'))
        d = np.load(os.path.join(HERE, 'c3_synth_all.npz'))
        off = np.concatenate([[0], np.cumsum(d['lens'])])
        sl = slice(off[k], off[k + 1])
        return dataclasses.replace(base, data=[d['wl'][sl], d['flux'][sl]], err=d['err'][sl])
    if name == 'c4':
        return synthetic.make_problem_c4(seed=2298)
    if name == 'c5':
        return synthetic.make_problem_c5(seed=2298)
    raise ValueError(name)


def thetas_for(prob, n_box, n_cloud, seed):
    from mcmc_spec_b200 import synthetic
    th = synthetic.sample_prior_box(prob, n_box, seed=seed)
    rng = np.random.default_rng(seed + 1000)
    centre = np.asarray(prob.truth, dtype=float).copy()
    cloud = centre * (1 + 0.02 * rng.standard_normal((n_cloud, centre.size)))
    cloud[:, 2] = np.abs(cloud[:, 2])
    return np.concatenate([th, cloud], axis=0)


_ORC = None


def _init(name, k):
    global _ORC
    os.environ['OMP_NUM_THREADS'] = '1'
    from oracle_adapter import to_oracle
    _ORC = to_oracle(problem_for(name, k))


def _eval(th):
    return float(_ORC.lnprob(th))


def run_case(name, theta, k=0):
    with mp.get_context('spawn').Pool(os.cpu_count(), initializer=_init, initargs=(name, k)) as pool:
        return np.array(pool.map(_eval, list(theta), chunksize=16))


def main():
    which = sys.argv[1:] or ['c1', 'c2', 'c3', 'c4', 'c5']
    for idx, name in enumerate(['c1', 'c2', 'c3', 'c4', 'c5']):
        if name not in which:
            continue
        t0 = time.time()
        if name == 'c3':
            per_box, per_cloud = 667, 167                           # 12 x 834 = 10 008 vectors
            thetas, lnps, pid = [], [], []
            for k in range(12):
                prob = problem_for('c3', k)
                th = thetas_for(prob, per_box, per_cloud, seed=300 + k)
                thetas.append(th); lnps.append(run_case('c3', th, k)); pid.append(np.full(len(th), k))
            theta, lnp = np.concatenate(thetas), np.concatenate(lnps)
            extra = {'problem': np.concatenate(pid)}
        else:
            prob = problem_for(name)
            theta = thetas_for(prob, N_BOX, N_CLOUD, seed=100 + idx)
            lnp = run_case(name, theta)
            extra = {}
        out = os.path.join(HERE, name + '_10k.npz')
        np.savez_compressed(out, theta=theta, oracle_lnprob=lnp, **extra)
        print(f"{name}: {len(theta)} vectors, {int(np.isfinite(lnp).sum())} finite, {time.time() - t0:.0f} s, "
              f"{os.path.getsize(out) // 1024} KiB", flush=True)


if __name__ == '__main__':
    main()
