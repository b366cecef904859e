#!/usr/bin/env python
"""bench.py -- walker log-likelihood evaluations/s on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config c4|c1|c2|c3|c5]      # product arm (CUDA, C ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W               # CPU arm (oracle port, all host cores)

Default workload (config C4; `--config c4`): a 65 536-walker ensemble on an 8192-pixel
observed spectrum, 32 768-pixel model rows, 26 x 4 (Teff, log g) grid, 2 contrast + 6 photometric bands, synthetic
tables (seed 2298).  One *step* is one ensemble iteration of the stretch move: two half steps, each proposing for half
of the walkers, evaluating `logposterior` for the proposals, accepting, and all-gathering the updated half over NCCL
when N > 1 -- i.e. 65 536 evaluations per step, sharded by walker across the N ranks (strong scaling: the ensemble
size is fixed by the config).  The other configs run the same loop on the other BASELINE shapes (C1 / C2: the
reference's own 50-walker ensemble on the committed real-data fixtures; C3: the 12 Data/synth_spec_* spectra as 12
independent 50-walker ensembles sharing one device grid; C5: 16 384 walkers on the 3-D 10^5-pixel grid).

`value`    whole-job evaluations/s, ensemble resident in HBM, timed with CUDA events around each step on the launching
           stream (L2 flushed between steps), max over ranks.
`e2e`      the same count through the public host-buffer call (`LogPosterior(theta_host)` -> ms_lnprob(MS_HOST)):
           host->device copy of the proposals and device->host copy of the log-probabilities inside the timed region.
`roofline` the evaluation kernel (event-timed inside the library): algorithmic bytes (n_star * 2^D * P_reg * 8 + 8 ndim
           + 8 per evaluation) per second against the measured HBM copy peak -- `bound`/`frac` as the
           contract words them -- and, because the grid is L2 resident and the kernel is arithmetic bound, `binding`:
           the fp64 work of the timed launches against the DFMA peak measured on this GPU model
           (profiles/sm_peaks.json, scripts/measure_sm_peaks.py); DRAM / L2 traffic per launch from the committed ncu
           capture (profiles/roofline_traffic.json).
`cpu_baseline` the numpy/scipy oracle (a port: the reference itself cannot be imported offline) on the host cores,
           bounded sample, rank 0 at N = 1 only.
`ensemble_checksum` exact fingerprint of the ensemble after the timed steps: identical at N = 1, 2, 4, 8 (the random
           streams and every walker's log-posterior are independent of the sharding).
"""
from __future__ import annotations

import argparse
import json
import math
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "walker log-likelihood evals/sec"
UNIT = "evals/s"
SEED = 2298

WORKLOADS = {
    'c4': dict(walkers=65536, text="C4: 65536-walker ensemble, 8192-px observed spectrum, 32768-px model rows, 26x4 grid, "
                                   "2 contrast + 6 photometric bands (synthetic, seed 2298)"),
    'c1': dict(walkers=50, text="C1: recovery of Data/synth_spec_3850_3425.txt with param_koi2298.txt settings (committed fixture: "
                                "observed spectrum + bps/ passbands), 135000-px rows, 26x4 synthetic grid, the reference's 50 walkers"),
    'c2': dict(walkers=50, text="C2: KOI-2298 HET spectrum + contrasts/photometry of param_koi2298.txt (committed fixture), "
                                "135000-px rows, 26x4 synthetic grid, the reference's 50 walkers"),
    'c3': dict(walkers=600, text="C3: the 12 Data/synth_spec_* spectra as 12 independent 50-walker ensembles sharing one device "
                                 "grid (committed fixture), one stream per ensemble"),
    'c5': dict(walkers=16384, text="C5: 16384-walker ensemble, 25000-px observed spectrum, 100000-px rows, 26x4x5 (Teff, log g, "
                                   "[Fe/H]) grid (416 MB), 4 contrast + 12 photometric bands (synthetic, seed 2298)"),
}


# ------------------------------------------------------------------------------------------ workload
def make_workload(name='c4', k=0):
    from mcmc_spec_b200 import synthetic
    if name == 'c4':
        return synthetic.make_problem_c4(seed=SEED)
    if name == 'c5':
        return synthetic.make_problem_c5(seed=SEED)
    import make_theta10k
    return make_theta10k.problem_for(name, k)


def initial_ensemble(prob, nw, seed=1, name='c4', init='ball'):
    """`ball` (default): a 1 % ball around the solution, the way the reference starts its walkers from the best third
    of the optimiser's results (mft6.py:3670-3674) -- the regime an ensemble sampler spends its life in.  `box`
    (--init box, C4 / C5 only): uniform in the central third of the prior box, i.e. a chain far from the mode whose
    walkers spread over the whole (Teff1, Teff2) plane; reported in DESIGN.md beside the default."""
    rng = np.random.default_rng(seed)
    if init == 'box' and name in ('c4', 'c5'):
        lo = np.array([min(prob.temps), min(prob.temps), 0.0, 0.1, 0.1, 1 / 3000.0])
        hi = np.array([max(prob.temps), max(prob.temps), 1.0, 1.2, 1.0, 1 / 100.0])
        a, b = lo + (hi - lo) / 3.0, hi - (hi - lo) / 3.0
        return a + (b - a) * rng.random((nw, 6))
    centre = np.asarray(prob.truth, dtype=float)
    p0 = centre * (1 + (0.01 if name in ('c4', 'c5') else 0.02) * rng.standard_normal((nw, centre.size)))
    p0[:, 2] = np.abs(p0[:, 2])
    return p0


def workload_config(name, prob, tables, nw, init='ball'):
    """Identical in the product arm and the reference arm (the driver compares the two)."""
    return {"workload": WORKLOADS[name]['text'], "name": name,
            "walkers": int(nw), "n_data": int(tables.n_data), "p_reg": int(tables.n_reg),
            "grid_nodes": ([len(prob.fehs)] if prob.fehs is not None else []) + [len(prob.lgs), len(prob.temps)],
            "bytes_per_eval": int(tables.bytes_per_eval()),
            "step": f"one stretch-move iteration = 2 half steps = {int(nw)} logposterior evaluations",
            "walker_init": ("uniform in the central third of the prior box (a chain far from the mode)" if init == 'box' and name in ('c4', 'c5')
                            else "ball around the solution (1 % for C4/C5, 2 % for C1-C3), as the reference starts from the "
                                 "optimiser's best third (mft6.py:3670-3674)"),
            "l2": "flushed between timed steps (256 MiB write); the grid rows are re-read by every group of walkers within a "
                  "step and are L2 resident by nature"}


# ------------------------------------------------------------------------------------------ CPU arm
_ORACLE = None


def _oracle_init(name):
    global _ORACLE
    os.environ['OMP_NUM_THREADS'] = '1'
    from oracle_adapter import to_oracle
    _ORACLE = to_oracle(make_workload(name))


def _oracle_eval(theta):
    return float(_ORACLE.lnprob(theta))


def cpu_pool(cores, name):
    ctx = mp.get_context('spawn')           # no fork after CUDA / torch threads
    return ctx.Pool(cores, initializer=_oracle_init, initargs=(name,))


def time_cpu_sample(pool, theta):
    t0 = time.perf_counter()
    out = pool.map(_oracle_eval, list(theta), chunksize=max(1, len(theta) // (4 * pool._processes)))
    return time.perf_counter() - t0, np.array(out)


def cpu_baseline(prob, name, target_s=15.0, init='ball'):
    """Oracle throughput on all host cores over a bounded sample of the workload's ensemble."""
    cores = os.cpu_count() or 1
    theta = initial_ensemble(prob, 65536, seed=2, name=name, init=init)
    with cpu_pool(cores, name) as pool:
        dt, _ = time_cpu_sample(pool, theta[:cores * 2])                  # warm-up + per-eval estimate
        dt, _ = time_cpu_sample(pool, theta[:cores * 2])
        per = dt / (cores * 2)
        n = int(min(len(theta), max(cores * 4, target_s / max(per, 1e-6))))
        dt, out = time_cpu_sample(pool, theta[:n])
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} evaluations of the {name.upper()} workload (walkers drawn like the GPU ensemble) through the numpy/scipy "
                      f"oracle, multiprocessing.Pool({cores}), {dt:.1f} s; optimistic w.r.t. the reference (no glob / "
                      f"pyphot / astropy / dustmaps overhead)"}


def run_reference_arm(args, rank):
    if rank != 0:
        return 0
    name = args.config
    prob = make_workload(name)
    from mcmc_spec_b200.tables import build_tables
    tables = build_tables(prob)
    nw = args.walkers or WORKLOADS[name]['walkers']
    cores = os.cpu_count() or 1
    per_step = cores * 8
    theta = initial_ensemble(prob, per_step * (args.steps + args.warmup), seed=3, name=name, init=args.init)
    times = []
    with cpu_pool(cores, name) as pool:
        time_cpu_sample(pool, theta[:cores])                                  # workers up
        for k in range(args.warmup + args.steps):
            dt, _ = time_cpu_sample(pool, theta[k * per_step:(k + 1) * per_step])
            if k >= args.warmup:
                times.append(dt)
    total = float(np.sum(times))
    value = per_step * args.steps / total
    sample = (f"each timed step = {per_step} logposterior evaluations (a bounded sample of the workload's "
              f"{nw}-evaluation step) through the numpy/scipy oracle on multiprocessing.Pool({cores})")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(name, prob, tables, nw, args.init),
            "reference_step": sample,
            "reference_note": "mft6.py cannot be imported offline (emcee, pyphot, extinction, dustmaps ... absent); this arm "
                              "times the oracle port (oracle/mft6_oracle.py)",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), [s.strip() for s in line.split(',')]))

    def wait_ready(self, timeout):
        t0 = time.time()
        while self.proc and not self.samples and time.time() - t0 < timeout:
            time.sleep(0.02)

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        rows = [r for t, r in self.samples if t0 <= t <= t1 + 0.15 and len(r) >= 7]
        if not rows:
            rows = [r for _, r in self.samples if len(r) >= 7][-3:]
        if not rows:
            return None
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k] == "Active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows)}


# ------------------------------------------------------------------------------------------ product arm
def _load_json(*parts):
    try:
        with open(os.path.join(ROOT, *parts)) as f:
            return json.load(f)
    except Exception:
        return None


def measured_peak():
    d = _load_json('MEASURED_PEAKS.json')
    if d and 'hbm_gbs' in d:
        return float(d['hbm_gbs']), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def fp64_ops_per_eval(tables):
    """fp64 instructions (DFMA / DMUL / DADD, one per lane) one evaluation needs in the group kernel, counted from
    the source: per observed pixel 16 (composite at the two bracketing pixels) + 2 (sum over
    the stars) + 15 (two extinction factors: exponent + 8th-order polynomial for the lower bracket, a 4th-order step to
    the upper one) + 5 (extinction, rebin, scale) + 3 (reciprocal) + 3 (moments) in the stream, 2 + 3 + 4 in the
    chi-square pass.  The rows' passband integrals and bracket pairs come from tables built once per handle."""
    return 53.0 * tables.n_data


def run_product_arm(args, rank, local_rank, world):
    name = args.config
    cpu = None
    prob = make_workload(name)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(prob, name, target_s=args.cpu_seconds, init=args.init)     # before CUDA is initialised

    import torch
    import torch.distributed as dist
    from mcmc_spec_b200.ensemble import CudaBackend, EnsembleSampler
    from mcmc_spec_b200.posterior import LogPosterior

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'            # keep NCCL's version banner off stdout (one JSON line)
        # NCCL prints its version banner on stdout when the first communicator comes up: keep stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if world > 1:
            dist.barrier()

    nw = args.walkers or WORKLOADS[name]['walkers']
    work_stream = torch.cuda.Stream(device=dev)                     # everything of the step runs on this stream
    torch.cuda.set_stream(work_stream)
    # C3: 12 independent problems (one LogPosterior each, one device copy of the grid, one stream each)
    n_prob = 12 if name == 'c3' else 1
    lps, samplers, streams = [], [], []
    for k in range(n_prob):
        p_k = prob if k == 0 else make_workload(name, k)
        lp_k = LogPosterior(p_k, device=local_rank, share_grid_with=(lps[0] if lps else None))
        lps.append(lp_k)
        s_k = EnsembleSampler(nw // n_prob, 6, CudaBackend(lp_k.engine), seed=SEED + k, rank=rank, world_size=world)
        samplers.append(s_k)
        streams.append(work_stream if n_prob == 1 else torch.cuda.Stream(device=dev))
    lp, tables = lps[0], lps[0].tables
    for k, s_k in enumerate(samplers):
        with torch.cuda.stream(streams[k]):
            s_k.initialize(initial_ensemble(prob, nw // n_prob, seed=1 + k, name=name, init=args.init))
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

    def step_all():
        if n_prob == 1:
            samplers[0].step()
            return
        ev0 = torch.cuda.Event(); ev0.record(work_stream)
        for k, s_k in enumerate(samplers):
            streams[k].wait_event(ev0)
            with torch.cuda.stream(streams[k]):
                s_k.step()
            e = torch.cuda.Event(); e.record(streams[k]); work_stream.wait_event(e)

    clocks = ClockSampler(local_rank) if rank == 0 else None        # nvidia-smi needs a moment to come up: start it before the warm-up
    for _ in range(args.warmup):
        step_all()
    torch.cuda.synchronize()
    if clocks:
        clocks.wait_ready(3.0)
    barrier()
    for l in lps:
        l.engine.set_timing(True); l.engine.kernel_time(); l.engine.short_circuited(reset=True)
    launches0 = sum(l.engine.launch_count for l in lps)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.time()
    for k in range(args.steps):
        flush.zero_()                                               # evict L2 between timed steps
        ev[k][0].record()
        step_all()
        ev[k][1].record()
    torch.cuda.synchronize(); barrier()
    t_wall1 = time.time()
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    total_ms = torch.tensor([float(step_ms.sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) / 1e3
    k_ms = k_launches = k_walkers = 0
    for l in lps:
        a, b, c = l.engine.kernel_time()
        k_ms += a; k_launches += b; k_walkers += c
        l.engine.set_timing(False)
    cut_short = sum(l.engine.short_circuited() for l in lps)
    launches = sum(l.engine.launch_count for l in lps) - launches0
    acc_frac = float(np.mean([s.acceptance_fraction.mean() for s in samplers]))
    finite_frac = float(np.mean([torch.isfinite(s.log_prob).double().mean().item() for s in samplers]))
    checksum = 0
    for s in samplers:
        checksum = (checksum + s.checksum()) & 0xFFFFFFFFFFFFFFFF
    # Clocks under load: nvidia-smi samples every 100 ms, a short timed region (K steps of a few ms) may see none.  The same
    # loop then goes on untimed (every rank the same number of steps: total_s is the max over the ranks) until 0.4 s of
    # load have been sampled; timers, counters and the checksum above are those of the K timed steps.
    t_wall2 = t_wall1
    if total_s < 0.4:
        n_extra = min(2000, int(math.ceil((0.4 - total_s) / max(total_s / args.steps, 1e-6))))
        for _ in range(n_extra):
            flush.zero_()
            step_all()
        torch.cuda.synchronize(); barrier()
        t_wall2 = time.time()
    clock_summary = None
    if clocks:
        time.sleep(0.12)
        clocks.stop()
        clock_summary = clocks.summary(t_wall0, t_wall2)
        if clock_summary is not None and t_wall2 > t_wall1:
            clock_summary["window"] = "the timed steps + an untimed continuation of the same loop (0.4 s of load)"

    # ---- end-to-end through the host-buffer API (pinned staging + H2D + kernels + D2H per call)
    lp.engine.use_own_stream()
    sampler = samplers[0]
    per_rank = sampler.nwalkers // world
    # the inputs are this rank's share of the ensemble as the timed chain left it (host copy): the same kind of
    # walkers the device-resident measurement evaluates
    host_theta = np.ascontiguousarray(sampler.coords.detach().cpu().numpy()[rank * per_rank:(rank + 1) * per_rank])
    halves = [np.ascontiguousarray(host_theta[:per_rank // 2]), np.ascontiguousarray(host_theta[per_rank // 2:])]
    for _ in range(2):
        for h in halves:
            lp(h)
    lp.engine.use_own_stream()
    torch.cuda.synchronize(); barrier()
    t0 = time.perf_counter()
    e2e_sum = 0.0
    for _ in range(args.e2e_steps):
        for h in halves:
            e2e_sum += float(lp(h)[0])
    torch.cuda.synchronize(); barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_nw = sampler.nwalkers
    e2e_value = e2e_nw * args.e2e_steps / float(e2e_s.item())

    if rank == 0:
        peak, peak_src = measured_peak()
        bpe = tables.bytes_per_eval()
        achieved = (k_walkers * bpe / 1e9) / (k_ms / 1e3) if k_ms > 0 else 0.0
        traffic = _load_json('profiles', 'roofline_traffic.json') or {}
        peaks = _load_json('profiles', 'sm_peaks.json') or {}
        shape_key = f"{name}_{int(round(k_walkers / max(k_launches, 1)))}"
        tr = traffic.get(shape_key) or traffic.get(name) or {}
        fp64_ops = fp64_ops_per_eval(tables)
        fp64_tflops = 2.0 * fp64_ops * k_walkers / (k_ms / 1e3) / 1e12 if k_ms > 0 else 0.0
        n_half = 2 * args.steps * n_prob
        roofline = {"bound": "hbm", "kernel": "lnprob_group_kernel" if lp.engine.group_supported and k_walkers / max(k_launches, 1) >= 1024
                    else "lnprob_stream_kernel",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                    "traffic": tr.get("dram_bytes_per_launch"),
                    "avg_launch_ms": k_ms / max(k_launches, 1), "launches_timed": int(k_launches),
                    "walkers_per_launch": k_walkers / max(k_launches, 1),
                    "algorithmic_bytes_per_launch": bpe * k_walkers / max(k_launches, 1),
                    "kernel_share_of_step": k_ms / float(step_ms.sum()),
                    "hbm_frac_of_measured_traffic": (tr["dram_bytes_per_launch"] / 1e9 / (k_ms / max(k_launches, 1) / 1e3) / peak
                                                     if tr.get("dram_bytes_per_launch") and k_launches else None),
                    "l2_traffic": tr.get("l2_read_bytes_per_launch"),
                    "binding": {"unit": "fp64 (DFMA pipe)", "achieved_tflops": fp64_tflops,
                                "peak_tflops": peaks.get("dfma_tflops_peak"),
                                "frac": (fp64_tflops / peaks["dfma_tflops_peak"]) if peaks.get("dfma_tflops_peak") else None,
                                "fp64_ops_per_eval": fp64_ops,
                                "peak_source": "profiles/sm_peaks.json (scripts/measure_sm_peaks.py on a B200 of this pool)",
                                "ncu": tr.get("ncu_summary")},
                    "note": "the grid is L2 resident (DRAM traffic per launch is the scratch round trip of the model vectors, "
                            "far below the algorithmic row bytes), so frac against the HBM copy peak only says how much the "
                            "on-chip reuse multiplies HBM; the unit that binds is the fp64 pipe (see binding)"}
        line = {"metric": METRIC, "value": nw * args.steps / total_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(name, prob, tables, nw, args.init),
                "parallelism": ("1 GPU, no collective" if world == 1 else
                                f"every rank holds the ensemble, evaluation sharded by 32-walker groups x{world}, 1 all-reduce of the log-probs per half step"
                                if getattr(sampler, '_by_group', False) else f"walker-sharded x{world}, 1 all-gather per half step"),
                "acceptance_fraction": acc_frac, "finite_lnprob_frac": finite_frac,
                "proposals_cut_short_by_prior": int(cut_short),
                "ensemble_checksum": f"{checksum:016x}",
                "half_step_us": {"total": 1e3 * float(step_ms.sum()) / n_half * n_prob,
                                 "evaluation_kernel": 1e3 * k_ms / max(k_launches, 1),
                                 "everything_else": 1e3 * (float(step_ms.sum()) / n_half * n_prob - k_ms / max(k_launches, 1)),
                                 "note": "everything_else = proposal/preparation, cell sort + group table, acceptance, "
                                         "all-gather + scatter (N > 1), launch gaps"},
                "roofline": roofline,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_nw * 6 * 8),
                        "d2h_bytes_per_step": int(e2e_nw * 8), "steps": args.e2e_steps,
                        "api": "LogPosterior.__call__(theta_host[nw/2/N, 6]) -> ms_lnprob(MS_HOST), two calls per step per rank; inputs = the ensemble after the timed chain"},
                "gpu_launches": int(launches), "clocks": clock_summary}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    for l in reversed(lps):
        l.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=60)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c4', choices=sorted(WORKLOADS))
    ap.add_argument('--init', default='ball', choices=['ball', 'box'])
    ap.add_argument('--walkers', type=int, default=0)
    ap.add_argument('--e2e-steps', type=int, default=10)
    ap.add_argument('--cpu-seconds', type=float, default=15.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        return run_reference_arm(args, rank)
    if world != args.gpus:
        if args.gpus > 1:
            raise SystemExit(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    return run_product_arm(args, rank, local_rank, world)


if __name__ == '__main__':
    sys.exit(main())
