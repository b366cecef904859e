#!/usr/bin/env python
"""Device-resident throughput of lnprob_stream_kernel on every BASELINE configuration shape.

Supplementary to bench.py (which measures C4 through the sampler): for each shape, `nw` walkers
uniform in the central third of the prior box are evaluated with ms_lnprob(MS_DEVICE); the
streaming kernel is timed by the library's own CUDA events.  Prints one JSON line per config:
evals/s, algorithmic GB/s (B_eval = n_star*2^D*P_reg*8 + 8*ndim + 8) and the fraction of the
measured HBM copy peak.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np   # noqa: E402
import torch         # noqa: E402

from mcmc_spec_b200 import synthetic                  # noqa: E402
from mcmc_spec_b200.posterior import LogPosterior      # noqa: E402


def central_third(prob, nw, seed=1):
    rng = np.random.default_rng(seed)
    lo = np.array([min(prob.temps), min(prob.temps), 0.0, 0.1, 0.1, 1 / 3000.0])
    hi = np.array([max(prob.temps), max(prob.temps), 1.0, 1.2, 1.0, 1 / 100.0])
    a, b = lo + (hi - lo) / 3.0, hi - (hi - lo) / 3.0
    th = a + (b - a) * rng.random((nw, 6))
    if getattr(prob, 'feh_free', False):
        # free metallicity (HBM-bound stress): the whole Teff and [Fe/H] node ranges, so that the walkers touch
        # every reachable row of the grid (~26 x 2-3 x 5 rows x 0.8 MB, far beyond the 126 MB L2)
        th[:, 0:2] = rng.uniform(lo[0], hi[0], (nw, 2))
        th = np.concatenate([th, rng.uniform(min(prob.fehs), max(prob.fehs), (nw, 1))], axis=1)
    return th


def run(name, prob, nw, reps=5, cell_order=None, fast=False):
    peak = 6455.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    lp = LogPosterior(prob)
    if os.environ.get('MS_PAIRING') == '0':
        lp.engine.set_pairing(False)
    lp.engine.set_cell_order(cell_order)
    lp.engine.set_fast(fast)
    th = torch.tensor(central_third(prob, nw), device='cuda')
    out = torch.empty(nw, dtype=torch.float64, device='cuda')
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device='cuda')
    lp.engine.set_stream(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        lp.engine.lnprob_device(th.data_ptr(), nw, out.data_ptr())
    torch.cuda.synchronize()
    lp.engine.set_timing(True); lp.engine.kernel_time()
    for _ in range(reps):
        flush.zero_()
        lp.engine.lnprob_device(th.data_ptr(), nw, out.data_ptr())
    torch.cuda.synchronize()
    ms, nl, nwt = lp.engine.kernel_time()
    t = lp.tables
    bpe = lp.tables.bytes_per_eval()
    if fast:          # bracket-only mode: 2 model pixels per observed pixel + the band table, per row
        ncorner = t.nspec * (4 if t.feh_nodes.size <= 1 else 8)
        bpe = ncorner * (2 * t.n_data + t.contrast.n + t.phot.n) * 8 + 8 * lp.ndim + 8
    gbs = nwt * bpe / 1e9 / (ms / 1e3)
    print(json.dumps({"config": name, "walkers_per_launch": nw, "n_data": t.n_data, "p_reg": t.n_reg,
                      "grid_mb": round(t.grid.nbytes / 1e6, 1), "bytes_per_eval": bpe,
                      "kernel_ms_per_launch": ms / nl, "evals_per_s": nwt / (ms / 1e3), "algorithmic_gbs": gbs,
                      "frac_of_measured_hbm_peak": gbs / peak,
                      "finite_frac": float(torch.isfinite(out).double().mean().item())}), flush=True)
    lp.close()
    del th, out, flush
    torch.cuda.empty_cache()


def run_c1_sampler(nw=50, steps=2000):
    """Latency of one stretch-move iteration of the reference-sized ensemble (50 walkers, C1/C2 shape):
    the one-call fused iteration against the per-kernel path driven from Python."""
    from mcmc_spec_b200.ensemble import CudaBackend, EnsembleSampler
    prob = synthetic.make_problem(truth=synthetic.consistent_truth(3850.0, 3425.0, 0.12, 2.0732e-3))
    lp = LogPosterior(prob)
    p0 = prob.truth * (1 + 0.01 * np.random.default_rng(0).standard_normal((nw, 6)))
    p0[:, 2] = np.abs(p0[:, 2])
    res = {}
    for name, fused in (('fused_one_call', True), ('per_kernel_calls', False)):
        s = EnsembleSampler(nw, 6, CudaBackend(lp.engine), seed=1)
        s.fused = fused
        s.initialize(p0)
        s.run_mcmc(50)
        torch.cuda.synchronize()
        lp.engine.set_timing(True); lp.engine.kernel_time()
        t0 = time.perf_counter()
        s.run_mcmc(steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ms, nl, nwt = lp.engine.kernel_time()
        lp.engine.set_timing(False)
        res[name] = {"ms_per_iteration": 1e3 * dt / steps, "evals_per_s": nw * steps / dt,
                     "stream_kernel_ms_per_iteration": ms / steps}
    print(json.dumps({"config": "C1 sampler: 50 walkers (25 proposals per half step), C1/C2 shape", **res}), flush=True)
    lp.close()


def run_c1_host_calls(nw=25, reps=2000):
    """The drop-in use of the reference's integration: emcee calls ``lnprob(P[25, 6])`` with host arrays once per
    half step (vectorize=True).  Latency of that call (pinned staging, H2D, prep + streaming kernel, D2H, sync)."""
    prob = synthetic.make_problem(truth=synthetic.consistent_truth(3850.0, 3425.0, 0.12, 2.0732e-3))
    lp = LogPosterior(prob)
    th = prob.truth * (1 + 0.01 * np.random.default_rng(0).standard_normal((nw, 6)))
    th[:, 2] = np.abs(th[:, 2])
    for _ in range(20):
        lp(th)
    t0 = time.perf_counter()
    for _ in range(reps):
        lp(th)
    dt = (time.perf_counter() - t0) / reps
    one = th[0].copy()
    t0 = time.perf_counter()
    for _ in range(reps):
        lp(one)
    dt1 = (time.perf_counter() - t0) / reps
    print(json.dumps({"config": "C1 host calls: LogPosterior(P[25, 6]) as emcee's vectorised log_prob_fn, C1/C2 shape",
                      "ms_per_call_25_walkers": dt * 1e3, "evals_per_s": nw / dt,
                      "ms_per_call_scalar_form": dt1 * 1e3}), flush=True)
    lp.close()


def run_c3(n_fits=12, nw=25, reps=50):
    """BASELINE config C3: 12 independent 50-walker ensembles (one half step = 25 proposals each) on one GPU,
    one device copy of the grid, every fit on its own stream so their launches overlap."""
    import dataclasses
    base = synthetic.make_problem()
    rng = np.random.default_rng(3)
    lps, owner = [], None
    for k in range(n_fits):
        # distinct observed spectra (here: the base spectrum with its own noise draw), same grid and region
        flux = base.data[1] * (1 + 0.01 * rng.standard_normal(base.n_data))
        prob = dataclasses.replace(base, data=[base.data[0], flux / np.median(flux)])
        lp = LogPosterior(prob, share_grid_with=owner)
        owner = owner or lp
        lps.append(lp)
    th = [torch.tensor(central_third(base, nw, seed=10 + k), device='cuda') for k in range(n_fits)]
    out = [torch.empty(nw, dtype=torch.float64, device='cuda') for _ in range(n_fits)]
    def round_():
        for lp, t, o in zip(lps, th, out):
            lp.engine.lnprob_device(t.data_ptr(), nw, o.data_ptr())       # each handle enqueues on its own stream
    for _ in range(5):
        round_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        round_()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    # the same work issued one fit after the other on a single stream
    s = torch.cuda.current_stream().cuda_stream
    for lp in lps:
        lp.engine.set_stream(s)
    round_(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        round_()
    torch.cuda.synchronize()
    dt1 = (time.perf_counter() - t0) / reps
    bpe = lps[0].tables.bytes_per_eval()
    print(json.dumps({"config": "C3: 12 independent fits x 25 proposals per half step, one shared grid (C1/C2 shape)",
                      "fits": n_fits, "walkers_per_launch": nw, "ms_per_round_concurrent_streams": dt * 1e3,
                      "ms_per_round_single_stream": dt1 * 1e3, "evals_per_s": n_fits * nw / dt,
                      "algorithmic_gbs": n_fits * nw * bpe / dt / 1e9,
                      "finite_frac": float(np.mean([torch.isfinite(o).double().mean().item() for o in out]))}), flush=True)
    for lp in reversed(lps):
        lp.close()


if __name__ == '__main__':
    which = sys.argv[1:] or ['c1', 'c1h', 'c1s', 'c3', 'c4', 'c5']
    if 'fast' in which:
        run('FAST C1/C2 shape (bracket-only mode), 8192 walkers', synthetic.make_problem(), 8192, fast=True)
        run('FAST C1/C2 shape, reference walker count (50)', synthetic.make_problem(), 50, reps=20, fast=True)
        run('FAST C4 (bracket-only mode), 32768 walkers', synthetic.make_problem_c4(), 32768, fast=True)
        run('FAST C5 fixed [Fe/H] (bracket-only mode), 16384 walkers', synthetic.make_problem_c5(), 16384, reps=3, fast=True)
    if 'c1' in which:
        run('C1/C2 shape: 26x4 grid, 135000-px rows (P_reg ~117k synthetic bands), N=1945', synthetic.make_problem(), 8192)
        run('C1/C2 shape, reference walker count (50)', synthetic.make_problem(), 50, reps=20)
    if 'c1h' in which:
        run_c1_host_calls()
    if 'c1s' in which:
        run_c1_sampler()
    if 'c3' in which:
        run_c3()
    if 'c4' in which:
        run('C4: 26x4 grid, P_reg=32768, N=8192', synthetic.make_problem_c4(), 32768)
    if 'c5' in which:
        t0 = time.time()
        p5 = synthetic.make_problem_c5()
        print(f'# C5 problem built in {time.time() - t0:.1f} s', flush=True)
        run('C5: 26x4x5 grid (416 MB), P_reg=100000, N=25000, 4+12 bands, fixed [Fe/H] (2 of 5 planes in use)', p5, 16384, reps=3)
        import dataclasses
        run('C5 free [Fe/H]: same grid, metallicity per walker (all 5 planes in use: HBM bound)',
            dataclasses.replace(p5, feh_free=True), 16384, reps=3)
        run('C5 free [Fe/H], walkers launched in the caller\'s (random) order: rows come from HBM',
            dataclasses.replace(p5, feh_free=True), 16384, reps=3, cell_order=False)
